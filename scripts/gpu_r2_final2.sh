#!/bin/bash
# Round-2 final binary on one B200 (after the exchange-fence change): tests, bench lines, launch list, ncu traffic capture.
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_gpu.log
tail -3 $O/r2_pytest_gpu.log
for c in 2 3 4 5; do
  python bench.py --config $c > $O/r2_bench_c$c.json 2> $O/r2_bench_c$c.err
  python -c "import json;d=json.loads(open('$O/r2_bench_c$c.json').read().strip().splitlines()[-1]);print('config $c: ms %.3f value %.3e e2e %.3e frac %.3f cpu %.3e'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value']))"
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -c 300 $O/r2_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_engine -s 3 -c 1 -o $O/r2_engine -f \
  python bench.py --T 200 --steps 1 --warmup 3 --no-cpu-baseline > $O/r2_ncu_engine.log 2>&1
ls -la $O/r2_engine.ncu-rep
