"""llpf_b200 — B200-native particle-filter hot path (host-side mirror of the reference API).

Import as `import llpf_b200` (see the shim `llpf_b200.py` at the repo root).  The numerical work
happens in `csrc/libllpf_b200.so` (hand-written sm_100a CUDA behind the C-ABI of include/llpf.h);
nothing here falls back to a CPU implementation.
"""
from . import _abi  # noqa: F401
from ._abi import LLPFError, load_library  # noqa: F401
from .filters import (  # noqa: F401
    AbstractParticleFilter, AdvancedParticleFilter, AuxiliaryParticleFilter, CudaDynamics, CudaLikelihood, CudaMeasurement,
    GaussianLikelihood,
    LinearDynamics, LinearMeasurement, MvNormal, ParticleFilter, ParticleFilteringSolution, QuadtankRK4,
    ResampleMetropolis, ResampleResidual, ResampleStratified, ResampleSystematic, ResamplingStrategy, ancestors, bins, connect_shards, correct,
    effective_particles, expweights, forward_trajectory, index, last_run_ms, launch_count, loglik, loglik_batch, logsumexp,
    mean_trajectory, mode_trajectory, num_particles, particles, predict, resample, reset, set_state,
    shard_blob, shouldresample, trajectory_statistics, smooth, smoothed_cov, smoothed_mean, smoothed_trajs, last_smooth_ms, state, update,
    weighted_mean, weights, xprev)
from .rbpf import KalmanFilter, RBMeasurementModel, RBPF, rb_particles, rbpf_source  # noqa: F401
from .enkf import (  # noqa: F401
    EnsembleKalmanFilter, KalmanFilteringSolution, enkf_correct, enkf_covariance, enkf_forward_trajectory, enkf_particles,
    enkf_predict, enkf_reset, enkf_state, enkf_update)
from .estimation import (  # noqa: F401
    Normal, Uniform, log_likelihood_fun, metropolis, metropolis_batched, metropolis_threaded, naive_sampler, set_model, weighted_cov,
    weighted_quantile)
