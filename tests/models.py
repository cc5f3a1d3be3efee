"""Test-side model helpers: the product workload specs plus their oracle twins."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from llpf_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402


class _LG(W.LGSpec):
    def oracle_model(self):
        return O.ModelArrays(self.nx, self.nu, self.ny, self.C, self.R1, self.R2, self.mu0, self.Sigma0,
                             A=self.A, B=self.B, dynamics=0)

    def oracle_filter(self, N, filter=0, **kw):
        kw.setdefault("particle_dtype", self.dtype)
        return O.OracleFilter(self.oracle_model(), N, filter=filter, **kw)


def lg_model(nx=4, nu=2, ny=2, seed=0, r1=1.0, r2=1.0):
    s = W.lg_spec(nx, nu, ny, seed, r1, r2)
    return _LG(**s.__dict__)


def lg_large_model(nx=64, nu=2, ny=58, seed=0, dtype=np.float32, r1_offdiag=0.0):
    """config 5 model (Float32 particles); r1_offdiag != 0 makes R1 non-diagonal (general Cholesky path)."""
    s = W.lg_large_spec(nx, nu, ny, seed, dtype)
    if r1_offdiag:
        R1 = np.eye(nx) + r1_offdiag * (np.ones((nx, nx)) - np.eye(nx))
        s.R1 = R1
    return _LG(**s.__dict__)


class _QT(W.QuadtankSpec):
    def oracle_model(self):
        return O.ModelArrays(4, 2, 2, self.C, self.R1, self.R2, self.x0, self.R1, dynamics=1,
                             dyn_params=self.p, t_switch=self.t_switch, a1_factor=self.a1_factor,
                             integ_Ts=self.Ts, supersample=self.supersample)

    def oracle_filter(self, N, filter=1, **kw):
        kw.setdefault("resample_threshold", 0.5)
        return O.OracleFilter(self.oracle_model(), N, filter=filter, **kw)


def quadtank_model(**kw):
    return _QT(**kw)


def ref_model_2state():
    """The reference's end-to-end test model, test/runtests.jl:245-262 (n=2, m=1, p=1)."""
    A = np.array([[0.97043, -0.097368], [0.09736, 0.970437]])
    B = np.array([[0.1], [0.0]])
    C_ = np.array([[0.0, 1.0]])
    return A, B, C_
