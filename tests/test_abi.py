"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol include/llpf.h
declares, the ctypes struct layouts equal the C ones, and without a GPU the product fails loudly
(no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "llpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llpf_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    import llpf_b200 as L
    lib = L.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/llpf.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", built]).decode()
    exported = set(re.findall(r" T (llpf_[a-z_0-9]+)", out))
    assert set(declared) <= exported


def test_struct_layouts_match_c(tmp_path, built):
    from llpf_b200 import _abi
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "llpf.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(llpf_model), sizeof(llpf_config), sizeof(llpf_run_outputs));
  printf("%zu %zu %zu %zu %zu\n", offsetof(llpf_model, A), offsetof(llpf_model, dyn_params),
         offsetof(llpf_model, t_switch), offsetof(llpf_model, integ_Ts), offsetof(llpf_model, supersample));
  printf("%zu %zu %zu %zu %zu\n", offsetof(llpf_config, filter), offsetof(llpf_config, resample_threshold),
         offsetof(llpf_config, seed), offsetof(llpf_config, scan_mode), offsetof(llpf_config, world));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split("\n")
    sizes = [int(v) for v in lines[0].split()]
    assert sizes == [C.sizeof(_abi.Model), C.sizeof(_abi.Config), C.sizeof(_abi.RunOutputs)]
    M, K = _abi.Model, _abi.Config
    assert [int(v) for v in lines[1].split()] == [M.A.offset, M.dyn_params.offset, M.t_switch.offset,
                                                  M.integ_Ts.offset, M.supersample.offset]
    assert [int(v) for v in lines[2].split()] == [K.filter.offset, K.resample_threshold.offset, K.seed.offset,
                                                  K.scan_mode.offset, K.world.offset]


def test_sass_is_sm100a_and_has_no_fallback_paths(built):
    out = subprocess.check_output(["cuobjdump", "-lelf", built]).decode()
    assert "sm_100a" in out
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "lowlevelparticlefilters.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text


def _have_gpu():
    import llpf_b200 as L
    n = C.c_int()
    rc = L.load_library().llpf_device_count(C.byref(n))
    return rc == 0 and n.value > 0


def test_no_gpu_means_loud_failure(built):
    """On a box without a CUDA device every constructor raises (status NO_DEVICE) — never a CPU path."""
    if _have_gpu():
        pytest.skip("a GPU is present")
    import llpf_b200 as L
    from models import lg_model
    with pytest.raises(L.LLPFError) as e:
        lg_model().particle_filter(128)
    assert e.value.code == L._abi.ERR_NO_DEVICE
    with pytest.raises(L.LLPFError):
        L.resample(L.ResampleSystematic, np.full(8, 0.125), 0.5)
    with pytest.raises(L.LLPFError):
        L.logsumexp(np.zeros(4))


def test_missing_library_raises(tmp_path):
    from llpf_b200 import _abi
    with pytest.raises(OSError):
        _abi.load_library(str(tmp_path / "nope.so"))


def test_descriptor_validation(built):
    import llpf_b200 as L
    with pytest.raises(TypeError):
        L.ParticleFilter(10, lambda x, u, p, t: x, L.LinearMeasurement(np.eye(2)), L.MvNormal(np.eye(2)),
                         L.MvNormal(np.eye(2)), L.MvNormal(np.zeros(2), np.eye(2)))
    d = L.MvNormal(np.array([[2.0, 0.1], [0.1, 1.0]]))
    assert len(d) == 2 and np.all(d.mu == 0)
    d = L.MvNormal(np.ones(3), 4.0)
    assert d.Sigma.shape == (3, 3) and d.Sigma[1, 1] == 4.0


def test_column_major_marshalling(built):
    from llpf_b200.filters import _ModelBuffers
    import llpf_b200 as L
    A = np.arange(16.0).reshape(4, 4)
    B = np.arange(8.0).reshape(4, 2)
    Cm = np.arange(8.0).reshape(2, 4) + 100
    mb = _ModelBuffers(L.LinearDynamics(A, B), Cm, np.eye(4), np.eye(2), L.MvNormal(np.zeros(4), np.eye(4)))
    m = mb.struct
    assert (m.nx, m.nu, m.ny) == (4, 2, 2)
    a = np.ctypeslib.as_array(m.A, shape=(16,))
    assert a[1] == A[1, 0] and a[4] == A[0, 1]           # column-major, like a Julia Matrix
    c = np.ctypeslib.as_array(m.C, shape=(8,))
    assert c[1] == Cm[1, 0] and c[2] == Cm[0, 1]
