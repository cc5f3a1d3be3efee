"""ctypes view of include/llpf.h — struct layouts, enums and the library loader.

The product library is `csrc/libllpf_b200.so` (hand-written sm_100a CUDA behind a C-ABI).
There is NO CPU fallback: if the library is missing or no CUDA device is present the calls
raise.  (The CPU oracle under /oracle is test infrastructure and is never imported here.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libllpf_b200.so")

# status codes
OK, ERR_BAD_ARG, ERR_CUDA, ERR_NONFINITE, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NOT_POSDEF = range(7)
# filter kinds
FILTER_PF, FILTER_ADVANCED, FILTER_AUX, FILTER_AUX_ADVANCED = range(4)
# resampling
RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED, RESAMPLE_RESIDUAL, RESAMPLE_METROPOLIS = range(4)
# scan mode
SCAN_FAST, SCAN_SERIAL = range(2)
# dynamics
DYN_LINEAR, DYN_QUADTANK_RK4, DYN_USER = range(3)
# particle element type
PARTICLE_F64, PARTICLE_F32 = range(2)
# time convention
TIME_FORWARD_TRAJECTORY, TIME_LOGLIK = range(2)

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


class Model(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("nu", C.c_int32), ("ny", C.c_int32), ("dynamics", C.c_int32),
        ("A", c_double_p), ("B", c_double_p), ("C", c_double_p),
        ("R1", c_double_p), ("R2", c_double_p), ("mu0", c_double_p), ("Sigma0", c_double_p),
        ("dyn_params", C.c_double * 8),
        ("t_switch", C.c_double), ("a1_factor", C.c_double), ("integ_Ts", C.c_double),
        ("supersample", C.c_int32), ("_pad", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("filter", C.c_int32), ("resampling", C.c_int32),
        ("resample_threshold", C.c_double), ("Ts", C.c_double), ("seed", C.c_uint64),
        ("scan_mode", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("particle_dtype", C.c_int32), ("single_block", C.c_int32),
        ("metropolis_steps", C.c_int32), ("_reserved", C.c_int32),
    ]


class RunOutputs(C.Structure):
    _fields_ = [
        ("ll_steps", c_double_p), ("ess_steps", c_double_p), ("resampled", c_int32_p),
        ("xhat", c_double_p), ("x_hist", c_double_p), ("w_hist", c_double_p), ("we_hist", c_double_p),
    ]


class HistStats(C.Structure):
    _fields_ = [
        ("xmean", c_double_p), ("xmode", c_double_p), ("xcov", c_double_p), ("q", c_double_p),
        ("nq", C.c_int32), ("_pad", C.c_int32), ("xquantile", c_double_p),
    ]


class LLPFError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"llpf status {code}: {msg}")


_lib = None


def load_library(path=None):
    """Load libllpf_b200.so and declare the prototypes of include/llpf.h. Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("LLPF_LIB_PATH") or LIB_PATH   # LLPF_LIB_PATH: tuning variants of the same library
    if not os.path.exists(p):
        raise OSError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the product path.")
    lib = C.CDLL(p)
    H = C.c_void_p
    dp, ip, i32p = c_double_p, c_int64_p, c_int32_p
    protos = {
        "llpf_create": [C.POINTER(Config), C.POINTER(Model), C.POINTER(H)],
        "llpf_destroy": [H],
        "llpf_device_count": [C.POINTER(C.c_int)],
        "llpf_set_model": [H, C.POINTER(Model)],
        "llpf_create_user": [C.POINTER(Config), C.POINTER(Model), C.c_char_p, dp, C.c_int32, C.POINTER(H)],
        "llpf_set_user_params": [H, dp, C.c_int32],
        "llpf_reset": [H, C.c_uint64],
        "llpf_correct": [H, dp, dp, C.c_double, dp],
        "llpf_predict": [H, dp, C.c_double],
        "llpf_predict_aux": [H, dp, dp, C.c_double],
        "llpf_update": [H, dp, dp, dp, C.c_double, dp],
        "llpf_run": [H, C.c_int64, dp, dp, C.c_int32, C.c_uint64, dp, C.POINTER(RunOutputs)],
        "llpf_run_stats": [H, C.c_int64, dp, dp, C.c_uint64, dp, C.POINTER(RunOutputs), C.POINTER(HistStats)],
        "llpf_run_batch": [C.c_int32, C.POINTER(H), C.c_int64, dp, dp, C.c_int32, C.POINTER(C.c_uint64), dp],
        "llpf_run_dev": [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64, dp, C.POINTER(RunOutputs)],
        "llpf_smooth": [H, C.c_int64, dp, dp, C.c_int64, C.c_uint64, dp, dp, C.POINTER(RunOutputs)],
        "llpf_smooth_history": [H, C.c_int64, dp, dp, dp, dp, C.c_int64, C.c_uint64, dp],
        "llpf_last_smooth_ms": [H, C.POINTER(C.c_float)],
        "llpf_num_particles": [H, ip],
        "llpf_local_particles": [H, ip, ip],
        "llpf_index": [H, ip],
        "llpf_get_particles": [H, dp],
        "llpf_get_xprev": [H, dp],
        "llpf_get_weights": [H, dp],
        "llpf_get_expweights": [H, dp],
        "llpf_get_ancestors": [H, ip],
        "llpf_get_bins": [H, dp],
        "llpf_set_state": [H, dp, dp, C.c_int64],
        "llpf_effective_particles": [H, dp],
        "llpf_shouldresample": [H, i32p],
        "llpf_weighted_mean": [H, dp],
        "llpf_resample_systematic": [C.c_int64, dp, C.c_double, C.c_int64, ip, dp, C.c_int32, C.c_int32],
        "llpf_resample_stratified": [C.c_int64, dp, dp, C.c_int64, ip, dp, C.c_int32, C.c_int32],
        "llpf_resample_residual": [C.c_int64, dp, dp, C.c_int64, ip, dp, C.c_int32, C.c_int32],
        "llpf_resample_metropolis": [C.c_int64, dp, C.c_int64, C.c_int32, C.c_uint64, ip, C.c_int32],
        "llpf_logsumexp": [C.c_int64, dp, dp, dp, C.c_int32],
        "llpf_enkf_set_inflation": [H, C.c_double],
        "llpf_enkf_reset": [H, C.c_uint64],
        "llpf_enkf_state": [H, dp, dp, ip],
        "llpf_enkf_predict": [H, dp, C.c_double],
        "llpf_enkf_correct": [H, dp, dp, C.c_double, dp, dp, dp, dp],
        "llpf_enkf_run": [H, C.c_int64, dp, dp, C.c_uint64, dp, dp, dp, dp, dp, dp, dp, dp, dp],
        "llpf_shard_blob_size": [C.POINTER(C.c_size_t)],
        "llpf_shard_export": [H, C.c_void_p],
        "llpf_shard_connect": [H, C.c_void_p],
        "llpf_launch_count": [H, ip],
        "llpf_last_run_ms": [H, C.POINTER(C.c_float)],
        "llpf_device_pointers": [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
    }
    for name, args in protos.items():
        if os.environ.get("LLPF_LIB_ALLOW_MISSING") and not hasattr(lib, name):
            continue             # A/B timing against an older build of the library (scripts/tune.py only)
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
    lib.llpf_last_error.argtypes = []
    lib.llpf_last_error.restype = C.c_char_p
    lib._llpf_symbols = sorted(list(protos) + ["llpf_last_error"])
    if path is None:
        _lib = lib
    return lib


def check(lib, code):
    if code != OK:
        msg = lib.llpf_last_error()
        raise LLPFError(code, msg.decode() if msg else "")
