"""Configuration holders with the reference's names (psvi/hypergrad/diff_optimizers.py:51-61,107-154).  On the B200
path the optimiser arithmetic itself runs inside psvi_mf_unroll (PSVI_ADAM_HYPERGRAD: u = b2*u + (1-b2)*g^2 + 1e-12,
denom = sqrt(u / (1 - b2^t)) + eps, reference :184-213) and the fixed-point map Phi(w) = w - step_size * grad inner(w)
(GradientDescent, :51-61) is applied through Hessian-vector kernels in hypergradients.py."""


class GradientDescent:
    def __init__(self, loss_f=None, step_size=1e-4, data_or_iter=None):
        self.loss_f, self.step_size, self.dim_mult = loss_f, step_size, 1


class DifferentiableAdam:
    def __init__(self, loss_f=None, step_size=1e-3, data_or_iter=None, betas=(0.9, 0.999), eps=1e-8, step_cnt=1):
        if tuple(betas) != (0.9, 0.999) or eps != 1e-8:
            raise NotImplementedError("the fused inner loop implements the reference's defaults (betas .9/.999, eps 1e-8)")
        self.loss_f, self.step_size, self.dim_mult, self.step_cnt = loss_f, step_size, 3, step_cnt
