"""Compiles one engine instantiation with NVRTC (no GPU needed: offline target sm_100a) to check that the engine headers
stay NVRTC-clean — the basis of the planned user-function hook: a filter whose dynamics /
measurement functions are device snippets supplied at run time would be compiled like this once per model."""
import os, sys, time
try:
    from cuda.bindings import nvrtc
except Exception:                      # older cuda-python layout
    from cuda import nvrtc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lowlevelparticlefilters.jl_b200", "csrc")
SRC = b'''
#include "llpf_engine.cuh"
namespace llpf {
template __global__ void k_engine<4, 2, 0, 0>(const __grid_constant__ EngineP, const __grid_constant__ ModelP<4, 2>);
}
'''


def compile_engine():
    err, prog = nvrtc.nvrtcCreateProgram(SRC, b"user_engine.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-default-device", b"-lineinfo",
            ("-I" + CSRC).encode(), b"-I/usr/local/cuda/include"]
    t0 = time.time()
    (rc,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    dt = time.time() - t0
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    _, nb = nvrtc.nvrtcGetCUBINSize(prog)
    return int(rc), dt, log.decode(errors="replace").strip("\x00 \n"), int(nb)


if __name__ == "__main__":
    rc, dt, log, nb = compile_engine()
    print(f"nvrtc rc={rc} in {dt:.1f} s, cubin {nb} bytes")
    if log:
        print(log[:4000])
    sys.exit(0 if rc == 0 and nb > 0 else 1)
