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


# a model no descriptor can express: x+ = [x0 + p0 sin(x1) + u0, p1 x1 + 0.1 cos(t)], y = x0^2 + e, e ~ N(0, p2)
USER_SRC = b'''
#define LLPF_USER_MODEL
#include "llpf_engine.cuh"
namespace llpf_user {
template <>
__device__ void dynamics<2>(double (&x)[2], const double* u, const double* p, double t) {
  const double x0 = x[0] + p[0] * sin(x[1]) + u[0];
  const double x1 = p[1] * x[1] + 0.1 * cos(t);
  x[0] = x0; x[1] = x1;
}
template <>
__device__ double loglik<2>(const double (&x)[2], const double* u, const double* y, const double* p, double t) {
  const double r = y[0] - x[0] * x[0];
  return -0.5 * r * r / p[2] - 0.5 * log(6.283185307179586 * p[2]);
}
}  // namespace llpf_user
namespace llpf {
template __global__ void k_engine<2, 1, 2, 0>(const __grid_constant__ EngineP, const __grid_constant__ ModelP<2, 1>);
}
'''


def compile_engine(src=None):
    err, prog = nvrtc.nvrtcCreateProgram(SRC if src is None else src, b"user_engine.cu", 0, [], [])
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
    ok = True
    for name, src in (("k_engine<4,2,0,0>", None), ("k_engine<2,1,2,0> + user model", USER_SRC)):
        rc, dt, log, nb = compile_engine(src)
        print(f"{name}: nvrtc rc={rc} in {dt:.1f} s, cubin {nb} bytes")
        if log:
            print(log[:4000])
        ok = ok and rc == 0 and nb > 0
    sys.exit(0 if ok else 1)
