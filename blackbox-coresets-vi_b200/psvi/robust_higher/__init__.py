"""`robust_higher` (the reference's edited copy of facebookresearch/higher, psvi/robust_higher/) has no counterpart
object model on the B200 path: the functional module, the differentiable optimiser tape and autograd's double backward
are replaced as a whole by psvi_mf_nested_step (csrc/psvi_mf_engine.cu) -- T unrolled robust-Adam steps
(optim.py:299-367 incl. the +1e-8 / v==0-mask semantics), the outer objective and a hand-written reverse sweep in one
launch.  The name is kept so that `from psvi.robust_higher import innerloop_ctx` resolves; calling it explains where
the functionality lives."""


def innerloop_ctx(model, opt, device=None, copy_initial_weights=True, override=None, track_higher_grads=True):
    raise NotImplementedError(
        "innerloop_ctx/diffopt.step (reference robust_higher/__init__.py:28-95) is fused away: use "
        "PSVI.nested_step (psvi.inference.psvi_classes) or the C entry point psvi_mf_nested_step")


def monkeypatch(module, device=None, copy_initial_weights=True, track_higher_grads=True):
    raise NotImplementedError("monkeypatch (reference robust_higher/patch.py:490-540) is fused away: see innerloop_ctx")
