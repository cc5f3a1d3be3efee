"""All five BASELINE.json configurations at full size on ONE B200 (configs 4 and 5 are quoted on 8 GPUs: here the whole
problem runs on a single GPU, the sharded numbers are in multi_gpu_timing.py), device time by CUDA events, next to the CPU
oracle on a bounded sample of the same workload (single core)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llpf_b200 as L
from llpf_b200 import workloads as W
from models import lg_large_model, lg_model, quadtank_model

def gpu(pf, u, y, fn=L.loglik, reps=3):
    best = 1e9
    for r in range(reps):
        d = fn(pf, u, y, epoch=r + 1)
        best = min(best, L.last_run_ms(pf))
    return best

def cpu(of, u, y, T_cpu):
    t0 = time.perf_counter()
    of.loglik(u[:T_cpu], y[:T_cpu], epoch=1)
    return time.perf_counter() - t0

def report(name, N, T, ms, cpu_s, N_cpu, T_cpu):
    g = N * T / ms / 1e6
    c = N_cpu * T_cpu / cpu_s / 1e6
    print(f"{name}: N={N} T={T}  GPU {ms:9.3f} ms = {g:8.3f} G particle-steps/s | CPU oracle (1 core, N={N_cpu}, T={T_cpu}) "
          f"{c:7.3f} M particle-steps/s | ratio {g * 1e3 / c:7.0f}x", flush=True)

# config 1: example_lineargaussian.jl (nx=2 as written), N=500, T=200, forward_trajectory with full history
s = lg_model(2, 2, 2, seed=0); T = 200
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
pf = s.particle_filter(500, seed=1)
ms = gpu(pf, u, y, fn=lambda p, a, b, epoch: L.forward_trajectory(p, a, b, epoch=epoch))
of = s.oracle_filter(500, seed=1); t0 = time.perf_counter(); of.forward_trajectory(u, y, epoch=1, history=True); c = time.perf_counter() - t0
report("config 1 (PF nx=2, full history)", 500, T, ms, c, 500, T)
# config 2
s = lg_model(4, 2, 2, seed=0); T = 1000
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
pf = s.particle_filter(1 << 20, seed=1); ms = gpu(pf, u, y)
report("config 2 (PF nx=4 f64)", 1 << 20, T, ms, cpu(s.oracle_filter(1 << 17, seed=1), u, y, 40), 1 << 17, 40)
del pf
# config 3: AdvancedParticleFilter, quadtank RK4 supersample 2, threshold 0.5
q = quadtank_model(); T = 2000
u = q.inputs(T); of = q.oracle_filter(256, seed=4); _, y = of.simulate(u, 9)
pf = q.advanced_filter(1 << 18, seed=4); ms = gpu(pf, u, y)
d = L.loglik(pf, u, y, epoch=1, details=True)
print(f"   config 3 resample fraction {d['resampled'].mean():.3f}, ll {d['ll']:.3f}")
report("config 3 (AdvancedPF quadtank)", 1 << 18, T, ms, cpu(q.oracle_filter(1 << 15, seed=4), u, y, 40), 1 << 15, 40)
del pf
# config 4: AuxiliaryParticleFilter, N=2^22 (whole problem on one GPU)
s = lg_model(4, 2, 2, seed=0); T = 1000
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
pf = s.aux_filter(1 << 22, seed=1); ms = gpu(pf, u, y, reps=2)
report("config 4 (APF nx=4, N=2^22 on 1 GPU)", 1 << 22, T, ms, cpu(s.oracle_filter(1 << 17, filter=2, seed=1), u, y, 30), 1 << 17, 30)
del pf
# config 5: 64 states, Float32 particles, N=2^20, T=500 (whole problem on one GPU)
s = lg_large_model(seed=0); T = 500
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
pf = s.particle_filter(1 << 20, seed=1); ms = gpu(pf, u, y, reps=2)
report("config 5 (PF nx=64 f32, N=2^20 on 1 GPU)", 1 << 20, T, ms, cpu(s.oracle_filter(1 << 12, seed=1), u, y, 20), 1 << 12, 20)
