#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the particle-filter hot path.

A "step" is one pass of the hot path over one batch of synthetic input: one full `loglik`-style trajectory
(reset! + T fused correct!/predict! steps).  metric = particle-steps/s = N*T / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]

--config selects the BASELINE.json configuration (default 2 = the headline the metric is quoted on):
  2  ParticleFilter, 4-state linear-Gaussian, N = 2^20 per GPU, T = 1000, Float64, threshold 0.1        (weak scaling)
  3  AdvancedParticleFilter, quadtank RK4 (examples/example_quadtank.jl), N = 2^18 per GPU, T = 2000       (weak scaling)
  4  AuxiliaryParticleFilter, 4-state LG, N = 2^22 in total sharded over the GPUs, T = 1000               (strong scaling)
  5  ParticleFilter, 64-state LG (test/test_large.jl regime), Float32 particles, N = 2^20 in total, T = 500 (strong scaling)

`value`   : device-timed (CUDA events on the filter's stream) with u,y already resident in HBM.
`e2e`     : the same metric through the public host API (`llpf_b200.loglik(pf, u, y)`) with pinned HOST buffers: H2D of
            u,y and D2H of the result are inside the timed region.
`roofline`: algorithmic bytes of the engine launch / its CUDA-event duration vs MEASURED_PEAKS.json;
            `traffic` = DRAM bytes per launch from the committed ncu capture IF that capture was taken on the kernel
            binary that is running now (SASS hash recorded next to it), else null.
`cpu_baseline`: the CPU oracle (a port of the reference's Julia loops; the reference itself cannot run here — no julia)
            timed on a bounded sample on this box's host cores.
--impl reference : the reference arm == that CPU port, single-threaded like the reference's ParticleFilter path
            (src/PFtypes.jl:107-139 have no @threads); same `config.workload` string as the product arm, `extrapolated: true`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# SURVEY §8d algorithmic bytes per particle-step: PF / AdvancedPF 3*nx*s + 16 (+ 32 on resample steps); APF 5*nx*s + 72
CONFIGS = {
    2: dict(kind="pf", nx=4, nu=2, ny=2, log2n=20, T=1000, thr=0.1, dtype="f64", scaling="weak", alg=112, alg_res=32,
            kernel="k_engine<4,2,0,0>", sass="k_engineILi4ELi2ELi0ELi0", cpu=(20, 6.0e6),
            name="ParticleFilter 4-state linear-Gaussian (nx=4,nu=2,ny=2), N=2^20 per GPU, T=1000, f64, systematic "
                 "resampling, threshold 0.1 (BASELINE config 2); loglik semantics"),
    3: dict(kind="adv", nx=4, nu=2, ny=2, log2n=18, T=2000, thr=0.5, dtype="f64", scaling="weak", alg=112, alg_res=32,
            kernel="k_engine<4,2,1,0>", sass="k_engineILi4ELi2ELi1ELi0", cpu=(18, 1.9e6),
            name="AdvancedParticleFilter quadtank RK4 (example_quadtank.jl: supersample 2, R1=0.1 I, R2=1e-4 I), N=2^18 per "
                 "GPU, T=2000, f64, systematic resampling, threshold 0.5 (BASELINE config 3); loglik semantics"),
    4: dict(kind="aux", nx=4, nu=2, ny=2, log2n=22, T=1000, thr=0.1, dtype="f64", scaling="strong", alg=232, alg_res=0,
            kernel="k_engine<4,2,0,0> (aux_step)", sass="k_engineILi4ELi2ELi0ELi0", cpu=(20, 3.2e6),
            name="AuxiliaryParticleFilter 4-state linear-Gaussian, N=2^22 in total (sharded over the GPUs), T=1000, f64, "
                 "systematic resampling every step (BASELINE config 4); loglik semantics"),
    5: dict(kind="wide", nx=64, nu=2, ny=58, log2n=20, T=500, thr=0.1, dtype="f32", scaling="strong", alg=784, alg_res=32,
            kernel="k_engine_wide", sass="k_engine_wide", cpu=(12, 2.6e4),
            name="ParticleFilter 64-state linear-Gaussian (nx=64,nu=2,ny=58; test_large.jl regime), Float32 particles / "
                 "Float64 weights, N=2^20 in total (sharded over the GPUs), T=500, systematic resampling, threshold 0.1 "
                 "(BASELINE config 5); loglik semantics"),
}


def workload(cfg, T):
    """(spec, u, y) of the configuration: seeded synthetic inputs, data simulated from the model itself."""
    from llpf_b200 import workloads as W
    c = CONFIGS[cfg]
    if c["kind"] == "adv":
        spec = W.QuadtankSpec()
        u = spec.inputs(T)
        _, y = W.simulate_quadtank(spec, u, seed=1)
        return spec, u, y
    spec = W.lg_large_spec(seed=0) if c["kind"] == "wide" else W.lg_spec(c["nx"], c["nu"], c["ny"], seed=0)
    u = np.random.default_rng(0).standard_normal((T, c["nu"]))
    _, y = W.simulate_lg(spec, u, seed=1)
    return spec, u, y


def make_filter(cfg, spec, N, **kw):
    c = CONFIGS[cfg]
    kw.setdefault("resample_threshold", c["thr"])
    if c["kind"] == "adv":
        return spec.advanced_filter(N, **kw)
    if c["kind"] == "aux":
        return spec.aux_filter(N, **kw)
    return spec.particle_filter(N, **kw)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.idx], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _oracle_filter(cfg, N):
    """The CPU oracle's twin of the configuration (test infrastructure: only the cpu_baseline / reference legs use it)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from models import lg_large_model, lg_model, quadtank_model
    c = CONFIGS[cfg]
    if c["kind"] == "adv":
        return quadtank_model().oracle_filter(N, seed=1, resample_threshold=c["thr"])
    if c["kind"] == "wide":
        return lg_large_model(seed=0).oracle_filter(N, seed=1, resample_threshold=c["thr"])
    return lg_model(c["nx"], c["nu"], c["ny"], seed=0).oracle_filter(N, filter=2 if c["kind"] == "aux" else 0, seed=1,
                                                                    resample_threshold=c["thr"])


def cpu_baseline(cfg=2, seconds_target=12.0):
    """The oracle port on one host core (faithful: the reference's ParticleFilter path is single-threaded), on a bounded
    sample of the configuration's workload: the first time steps at the N stated in `sample` (full per-GPU N for configs
    2 and 3, 2^20 for config 4, 2^12 for the 64-state config 5), sized to take about `seconds_target` seconds."""
    c = CONFIGS[cfg]
    log2n_cpu, rate_guess = c["cpu"]
    N = 1 << log2n_cpu
    T_cpu = max(4, min(c["T"], int(seconds_target * rate_guess / N)))
    _, u, y = workload(cfg, T_cpu)
    of = _oracle_filter(cfg, N)
    t0 = time.perf_counter()
    r = of.loglik(u, y, epoch=1)
    dt = time.perf_counter() - t0
    return {"value": N * T_cpu / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"N=2^{log2n_cpu}, first T={T_cpu} steps of the workload, {dt:.1f} s, ll={r['ll']:.6f}",
            "host_cores_available": os.cpu_count()}


def _replica(args):
    return cpu_baseline(*args)


def cpu_replicas_all_cores(cfg=2, seconds_target=6.0):
    """Every host core runs its own single-threaded copy of the workload (independent filters, the pattern of the
    reference's metropolis_threaded, smoothing.jl:335-347): the most the reference's ParticleFilter path — which has no
    threading of its own (PFtypes.jl:107-139) — can get out of the box.  Aggregate particle-steps/s."""
    import multiprocessing as mp
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    C = max(1, min(avail, 32))          # each replica holds up to a 2^20-particle filter (~150 MB): bound the footprint
    t0 = time.perf_counter()
    try:
        with mp.get_context("fork").Pool(C) as pool:
            rs = pool.map(_replica, [(cfg, seconds_target)] * C)
    except Exception as e:  # noqa: BLE001  (the aggregate is extra information: never fail the arm over it)
        return {"value": None, "unit": "particle-steps/s", "cores": C, "kind": "port", "sample": f"failed: {e}"}
    wall = time.perf_counter() - t0
    return {"value": sum(r["value"] for r in rs), "unit": "particle-steps/s", "cores": C, "kind": "port",
            "sample": f"{C} independent single-threaded filters, each: {rs[0]['sample']}; wall {wall:.1f} s"}


def config_block(cfg, world):
    """The `config` object shared verbatim by both arms (same workload string => the driver's same_config check)."""
    c = CONFIGS[cfg]
    return {"workload": c["name"], "baseline_config": cfg, "n_gpus": world}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm for the path (oracle port), rank 0 only."""
    if rank != 0:
        return
    cfg = args.config
    c = CONFIGS[cfg]
    res = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_baseline(cfg, 2.0)
    for _ in range(args.steps):
        res.append(cpu_baseline(cfg, 8.0))
    v = statistics.mean(r["value"] for r in res)
    cb = dict(res[-1]); cb["value"] = v
    world = max(1, args.gpus)
    N_global = (1 << c["log2n"]) * (world if c["scaling"] == "weak" else 1)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * N_global * c["T"] / v, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": c["dtype"], "data": "synthetic",
        "config": config_block(cfg, world),
        # each timed step is a bounded sample (about 8 s of CPU work, see cpu_baseline.sample) of the workload on the
        # single-threaded CPU port of the reference loops; value = its rate, ms_per_step = that rate scaled to the full N*T
        "extrapolated": True,
        "cpu_baseline": cb,
        # one filter cannot use more than one thread in the reference; all cores only help independent filters:
        "cpu_replicas_all_cores": cpu_replicas_all_cores(cfg),
        "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def build_info():
    try:
        return json.load(open(os.path.join(ROOT, "lowlevelparticlefilters.jl_b200", "csrc", "build_info.json")))
    except Exception:  # noqa: BLE001
        return {}


def measured_traffic(cfg, T):
    """DRAM bytes per launch from the committed ncu capture of this configuration's kernel — only if the capture was taken on
    the SASS that is running now (hash recorded by build() in build_info.json), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t["configs"][str(cfg)]
        now = build_info().get("sass", {}).get(CONFIGS[cfg]["sass"], {}).get("sha16")
        if now is None or now != e["sass_sha16"]:
            return None, f"capture {e.get('source')} is of another kernel binary ({e['sass_sha16']} != {now})"
        n_local = e["n_local"]
        return e["dram_bytes_per_particle_step"] * n_local * T, e.get("source")
    except Exception as ex:  # noqa: BLE001
        return None, f"no capture ({ex})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--log2n", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = args.config
    c = dict(CONFIGS[cfg])
    if args.log2n is not None:
        c["log2n"] = args.log2n

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import llpf_b200 as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    weak = c["scaling"] == "weak"
    N_global = (1 << c["log2n"]) * (world if weak else 1)
    n_local = N_global // world
    T = args.T or c["T"]
    spec, u, y = workload(cfg, T)
    # ONE global filter; with N > 1 its particles are block-partitioned over the ranks and the kernels exchange
    # (max, sum exp, sum exp^2) partials, CDF offsets and resampled particles over NVLink peer memory
    # (torch.distributed only carries the IPC descriptors at set-up)
    pf = make_filter(cfg, spec, N_global, seed=1, device=local_rank, rank=rank, world=world)
    if world > 1:
        L.connect_shards(pf)
    par = "single GPU" if world == 1 else (f"particles sharded over {world} GPUs (dp{world}); in-kernel exchange over "
                                           "NVLink peer memory: 1 all-gather of the weight statistics per step; on "
                                           "resample steps the CDF totals and one packed particle exchange")

    # device-resident inputs for `value`
    u_dev = torch.from_numpy(u).cuda()
    y_dev = torch.from_numpy(y).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # pinned host inputs for `e2e`
    u_pin = torch.from_numpy(u).pin_memory().numpy()
    y_pin = torch.from_numpy(y).pin_memory().numpy()

    import ctypes as C
    lib = pf._lib
    ll = C.c_double()

    def run_dev(epoch):
        L._abi.check(lib, lib.llpf_run_dev(pf._h, T, C.c_void_p(u_dev.data_ptr()), C.c_void_p(y_dev.data_ptr()),
                                           L._abi.TIME_LOGLIK, epoch, C.byref(ll), None))
        return L.last_run_ms(pf)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for w in range(args.warmup):
        if world > 1:
            dist.barrier()
        run_dev(100 + w)
    if world > 1:
        dist.barrier()
    rho_probe = L.loglik(pf, u, y, epoch=99, details=True)
    rho = float(rho_probe["resampled"].mean())

    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    launches0 = L.launch_count(pf)
    kernel_ms = []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed trajectories (untimed)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()                 # ranks enter the launch together (a late peer would be waited for in-kernel)
        kernel_ms.append(run_dev(200 + k))
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = L.launch_count(pf) - launches0
    # whole step on the device = reset! kernel + engine launch; the engine is >99.9 % of it, so the
    # step time is taken as the CUDA-event time of the engine launch plus the measured reset! time
    t0 = time.perf_counter(); L.reset(pf, 1); reset_ms = (time.perf_counter() - t0) * 1e3
    step_ms_local = statistics.mean(kernel_ms)
    # e2e through the public API with pinned host buffers
    sync_all()
    e2e_t0 = time.perf_counter()
    for k in range(args.steps):
        if world > 1:
            dist.barrier()
        L.loglik(pf, u_pin, y_pin, epoch=300 + k)
    torch.cuda.synchronize()
    e2e_ms_local = (time.perf_counter() - e2e_t0) * 1e3 / args.steps
    clocks = sampler.stop()

    if world > 1:
        tt = torch.tensor([step_ms_local, e2e_ms_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, e2e_ms = tt.tolist()
    else:
        step_ms, e2e_ms = step_ms_local, e2e_ms_local

    units = float(N_global) * T
    value = units / (step_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    alg_per = c["alg"] + c["alg_res"] * rho
    alg_bytes_launch = float(n_local) * T * alg_per
    achieved = alg_bytes_launch / (step_ms_local * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(cfg, T) if n_local == (1 << CONFIGS[cfg]["log2n"]) else (None, "other size")

    if rank == 0:
        cblock = config_block(cfg, world)
        cblock.update({
            "global_particles": N_global, "particles_per_gpu": n_local, "T": T, "parallelism": par,
            "l2": "L2 flushed (256 MiB write) between timed trajectories; inside a trajectory the particle state is "
                  "re-read every time step by construction (sequential time loop)",
            "resample_fraction": rho,
            "timing": "CUDA events on the filter's stream around each engine launch (one launch = one trajectory)",
            "reset_ms_host_timed": reset_ms, "wall_s_timed_region": t_wall,
        })
        bi = build_info()
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": c["dtype"], "data": "synthetic",
            "config": cblock,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int((u.size + y.size) * 8), "d2h_bytes_per_step": 8 + 200},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle_step": alg_per,
                         "kernel": f"{c['kernel']} (persistent cooperative; 1 launch = N*T particle-steps)",
                         "kernel_ms": step_ms_local,
                         "kernel_sass_sha16": bi.get("sass", {}).get(c["sass"], {}).get("sha16")},
            "build": {k: bi.get(k) for k in ("source_sha16", "nvcc", "built_at", "host")},
        }
        if c["kind"] == "wide":
            # config 5 is bounded by FP32 FMA issue, not HBM: 2*nx*(nx+ny) + noise matvec flops per particle-step
            flops = 2.0 * c["nx"] * (c["nx"] + c["ny"]) + 2.0 * c["nx"]
            tf = flops * n_local * T / (step_ms_local * 1e-3) / 1e12
            line["roofline_fp32"] = {"bound": "fp32-fma", "achieved": tf, "peak": 74.0, "unit": "TFLOP/s", "frac": tf / 74.0,
                                     "flops_per_particle_step": flops,
                                     "peak_source": "148 SMs x 128 FMA/clk x 2 x 1.965 GHz (CUDA cores; tensor cores unused: "
                                                    "f32 parity with the reference's sgemv)"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
