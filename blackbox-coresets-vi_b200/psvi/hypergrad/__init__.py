"""Implicit hypergradients for `--trainer hyper` on the B200 path (replaces the reference's vendored hypertorch,
psvi/hypergrad/): the Jacobian products of the fixed-point map are fused Hessian-vector kernels, see hypergradients.py."""
from psvi.hypergrad.hypergradients import CG_normaleq, fixed_point, cg_normaleq_native, fixed_point_native  # noqa: F401
from psvi.hypergrad.diff_optimizers import DifferentiableAdam, GradientDescent  # noqa: F401
