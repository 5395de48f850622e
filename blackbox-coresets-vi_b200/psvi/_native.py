"""ctypes binding of libpsvi_b200.so (include/psvi_b200.h) -- the only door from Python to the CUDA hot path.

Thin on purpose: every function passes `tensor.data_ptr()`s, sizes and the current CUDA stream.  There is no CPU
fallback and no other backend: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libpsvi_b200.so")

MAX_LAYERS = 6
VMODE_IDENTITY, VMODE_SOFTMAX, VMODE_EXPALPHA_SOFTMAX = 0, 1, 2
ADAM_ROBUST_HIGHER, ADAM_TORCH, ADAM_HYPERGRAD = 0, 1, 2
NOISE_EXTERNAL, NOISE_PHILOX = 0, 1
PHASE_UNROLL, PHASE_REVERSE = 1, 2


class NativeError(RuntimeError):
    code = None


ERR_UNSUPPORTED = -2


class MfModel(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)), ("mc_samples", C.c_int32)]


class Noise(C.Structure):
    _fields_ = [("mode", C.c_int32), ("eps", C.c_void_p), ("seed", C.c_uint64), ("domain", C.c_uint32)]


_lib = None
_LAUNCHES = 0      # kernels of libpsvi_b200 launched by this process (bench.py reports it as gpu_launches)
EVENT_HOOK = None  # (start_event, end_event, stream): bench.py brackets the dominant kernel with CUDA events


def launch_count():
    return _LAUNCHES


def _count(n):
    global _LAUNCHES
    _LAUNCHES += n

_SIGS = {
    "psvi_last_error": (C.c_char_p, []),
    "psvi_device_sm_count": (C.c_int, []),
    "psvi_mf_num_theta": (C.c_int64, [C.POINTER(MfModel)]),
    "psvi_mf_traj_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int32]),
    "psvi_mf_gout_floats": (C.c_int64, [C.POINTER(MfModel), C.c_int32]),
    "psvi_mf_nested_step": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "psvi_mf_unroll": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                 C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]),
    "psvi_mf_outer_grad": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_float, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_inner_grad": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "psvi_mf_inner_hvp": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_eval_scratch_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int32, C.c_int32]),
    "psvi_mf_evaluate": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "psvi_mf_forward": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_lr_predictive_tc_scratch_bytes": (C.c_size_t, [C.POINTER(MfModel)]),
    "psvi_lr_predictive_tc": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "psvi_fn_tc_scratch_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int64, C.c_int32]),
    "psvi_fn_predictive_tc": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_fn_nll_tc": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "psvi_predictive_tc_slabs": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                           C.c_float, C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_fn_data_grad_tc_scratch_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int64]),
    "psvi_fn_data_grad_tc": (C.c_int, [C.POINTER(MfModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_stream_workspace_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int32]),
    "psvi_mf_unroll_stream": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                        C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_evaluate_stream": (C.c_int, [C.POINTER(MfModel), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_net_pass": (C.c_int, [C.POINTER(MfModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_net_pass_gaussian": (C.c_int, [C.POINTER(MfModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "psvi_net_pass_bernoulli": (C.c_int, [C.POINTER(MfModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_net_predict": (C.c_int, [C.POINTER(MfModel), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_fc_matvec": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_int32, C.c_void_p]),
    "psvi_fc_outer": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "psvi_fnl_workspace_bytes": (C.c_size_t, [C.POINTER(MfModel), C.c_int32, C.c_int32]),
    "psvi_fnl_pass": (C.c_int, [C.POINTER(MfModel), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_sample": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_tangent": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_reparam_grad": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "psvi_mf_reparam_hvp": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_mf_nkl_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "psvi_mf_nkl_kl": (C.c_int, [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "psvi_fc_sample": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                 C.c_void_p]),
    "psvi_fc_reparam_grad": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p]),
    "psvi_fc_reparam_hvp": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "psvi_adam_unroll_step": (C.c_int, [C.c_int64, C.c_float, C.c_float] + [C.c_void_p] * 8),
    "psvi_adam_unroll_reverse": (C.c_int, [C.c_int64, C.c_float, C.c_float] + [C.c_void_p] * 8),
    "psvi_lenet_num_theta": (C.c_int64, []),
    "psvi_lenet_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "psvi_lenet_pass": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_logits_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "psvi_philox_normal": (C.c_int, [C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p]),
}


def exported_symbols():
    """Every symbol include/psvi_b200.h declares (used by the CPU-side loader test)."""
    return sorted(_SIGS)


def lib():
    """Loads libpsvi_b200.so (built by `__graft_entry__.build()` / `make -C csrc`).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                              " -- there is no CPU/PyTorch fallback for the PSVI hot path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError("the PSVI hot path needs a CUDA device (sm_100a); there is no CPU fallback")
    n = lib().psvi_device_sm_count()
    if n <= 0:
        raise NativeError("libpsvi_b200: " + last_error())
    return n


def last_error():
    return lib().psvi_last_error().decode()


def _check(rc):
    if rc != 0:
        e = NativeError(f"libpsvi_b200 error {rc}: {last_error()}")
        e.code = rc
        raise e


def make_model(dims, mc_samples):
    dims = [int(d) for d in dims]
    if not (2 <= len(dims) <= MAX_LAYERS + 1):
        raise NativeError(f"unsupported number of layers: dims={dims}")
    m = MfModel()
    m.n_layers = len(dims) - 1
    for i, d in enumerate(dims):
        m.dims[i] = d
    m.mc_samples = int(mc_samples)
    return m


def make_noise(eps=None, seed=0, domain=0):
    n = Noise()
    if eps is not None:
        _chk(eps, torch.float32)
        n.mode, n.eps = NOISE_EXTERNAL, eps.data_ptr()
        n._keepalive = eps  # the struct only carries the raw pointer: pin the tensor's lifetime to it
    else:
        n.mode, n.eps = NOISE_PHILOX, None
    n.seed, n.domain = int(seed) & (2**64 - 1), int(domain) & 0xFFFFFFFF
    return n


def _chk(t, dtype):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise NativeError(f"expected a contiguous CUDA {dtype} tensor, got {t.device} {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def _p(t, dtype=torch.float32):
    return _chk(t, dtype)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def num_theta(model):
    return int(lib().psvi_mf_num_theta(C.byref(model)))


def traj_floats(model, T):
    return int(lib().psvi_mf_traj_bytes(C.byref(model), T)) // 4


def gout_floats(model, M):
    return int(lib().psvi_mf_gout_floats(C.byref(model), M))


def nested_step(model, noise, mu, rho, u, z, v, xb, yb, n_total_rows, N, vmode, alpha, T, lr, pseudo_scale, phase_mask,
                traj, gout, u_grad, v_grad, alpha_grad, loss_out, inner_losses):
    B = 0 if xb is None else xb.shape[0]
    _count(1)
    hook = EVENT_HOOK
    if hook is not None:
        hook[0].record(hook[2])
    _check(lib().psvi_mf_nested_step(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v),
                                     u.shape[0], _p(xb), _p(yb, torch.int32), B, n_total_rows, N, vmode, alpha, T, lr,
                                     pseudo_scale, phase_mask, _p(traj), _p(gout), _p(u_grad), _p(v_grad),
                                     _p(alpha_grad), _p(loss_out), _p(inner_losses), _stream()))
    if hook is not None:
        hook[1].record(hook[2])


def coreset_weights(v, N, vmode, alpha):
    """a = N f(v) on the device (host-side plumbing for the streaming entry points)."""
    if vmode == VMODE_IDENTITY:
        return (N * v).contiguous()
    sc = N * (float(torch.exp(torch.tensor(alpha))) if vmode == VMODE_EXPALPHA_SOFTMAX else 1.0)
    return (sc * torch.softmax(v, 0)).contiguous()


_stream_ws = {}


def _workspace(model, n_rows, device):
    n = (int(lib().psvi_mf_stream_workspace_bytes(C.byref(model), n_rows)) + 3) // 4
    key = (device, _stream())      # chains on different CUDA streams must not share scratch
    t = _stream_ws.get(key)
    if t is None or t.numel() < n:
        t = torch.zeros(n, device=device, dtype=torch.float32)
        _stream_ws[key] = t
    return t


def unroll(model, noise, mu, rho, adam_m, adam_v, step0, x, y, row_weights, v, N, vmode, alpha, T, lr, adam_mode, losses):
    """psvi_mf_unroll; models too large for the shared-memory-resident engine go through psvi_mf_unroll_stream."""
    _count(1)
    rc = lib().psvi_mf_unroll(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(adam_m), _p(adam_v), step0, _p(x),
                              _p(y, torch.int32), _p(row_weights), _p(v), x.shape[0], N, vmode, alpha, T, lr,
                              adam_mode, _p(losses), _stream())
    if rc != ERR_UNSUPPORTED:
        return _check(rc)
    _count(4 * T - 1)
    if row_weights is None:
        row_weights = coreset_weights(v, N, vmode, alpha)
    if adam_m is None:
        adam_m, adam_v = torch.zeros(2 * mu.numel(), device=mu.device), torch.zeros(2 * mu.numel(), device=mu.device)
    ws = _workspace(model, 0, mu.device)
    _check(lib().psvi_mf_unroll_stream(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(adam_m), _p(adam_v), step0,
                                       _p(x), _p(y, torch.int32), _p(row_weights), 0.0, x.shape[0], T, lr, adam_mode,
                                       _p(losses), _p(ws), _stream()))


def outer_grad(model, noise, mu, rho, u, z, v, xb, yb, n_total_rows, N, vmode, alpha, pseudo_scale, gout, u_grad, v_grad,
               alpha_grad, loss_out):
    B = 0 if xb is None else xb.shape[0]
    _count(1)
    _check(lib().psvi_mf_outer_grad(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v),
                                    u.shape[0], _p(xb), _p(yb, torch.int32), B, n_total_rows, N, vmode, alpha,
                                    pseudo_scale, _p(gout), _p(u_grad), _p(v_grad), _p(alpha_grad), _p(loss_out),
                                    _stream()))


def inner_grad(model, noise, mu, rho, u, z, v, N, vmode, alpha, grad, value):
    _count(1)
    _check(lib().psvi_mf_inner_grad(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v),
                                    u.shape[0], N, vmode, alpha, _p(grad), _p(value), _stream()))


def inner_hvp(model, noise, mu, rho, u, z, v, N, vmode, alpha, gdot, h_phi, h_u, h_v, h_alpha=None):
    _count(1)
    _check(lib().psvi_mf_inner_hvp(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v),
                                   u.shape[0], N, vmode, alpha, _p(gdot), _p(h_phi), _p(h_u), _p(h_v), _p(h_alpha),
                                   _stream()))


def eval_scratch_floats(model, n_rows, batch):
    return int(lib().psvi_mf_eval_scratch_bytes(C.byref(model), n_rows, batch)) // 4


def evaluate(model, noise, mu, rho, u, z, v, xt, yt, batch, first_slab, N, vmode, alpha, mode, out, scratch):
    """psvi_mf_evaluate; models too large for the shared-memory-resident kernels go through psvi_mf_evaluate_stream."""
    M = 0 if u is None else u.shape[0]
    _count(3 if mode == 0 else 2)
    rc = lib().psvi_mf_evaluate(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v), M,
                                _p(xt), _p(yt, torch.int32), xt.shape[0], batch, first_slab, N, vmode, alpha, mode,
                                _p(out), _p(scratch), _stream())
    if rc != ERR_UNSUPPORTED:
        return _check(rc)
    a = coreset_weights(v, N, vmode, alpha) if mode == 0 else None
    ws = _workspace(model, xt.shape[0], mu.device)
    _check(lib().psvi_mf_evaluate_stream(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(a),
                                         M, _p(xt), _p(yt, torch.int32), xt.shape[0], batch, first_slab, mode, _p(out),
                                         _p(ws), _stream()))


def philox_normal(seed, domain, first_slab, n_slabs, S, P, out):
    _count(1)
    _check(lib().psvi_philox_normal(int(seed) & (2**64 - 1), int(domain) & 0xFFFFFFFF, first_slab, n_slabs, S, P,
                                    _p(out), _stream()))


def forward(model, noise, mu, rho, x, logits, theta_out=None, nkl_out=None, kl_out=None):
    _count(1)
    _check(lib().psvi_mf_forward(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(x), x.shape[0], _p(logits),
                                 _p(theta_out), _p(nkl_out), _p(kl_out), _stream()))


def lr_predictive_tc_scratch_floats(model):
    return (int(lib().psvi_lr_predictive_tc_scratch_bytes(C.byref(model))) + 3) // 4


def lr_predictive_tc(model, noise, mu, rho, u, z, v, xt_bf16, yt, slab, N, vmode, alpha, mode, out, scratch):
    M = 0 if u is None else u.shape[0]
    _count(4 if mode == 0 else 3)
    _check(lib().psvi_lr_predictive_tc(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v), M,
                                       _p(xt_bf16, torch.bfloat16), _p(yt, torch.int32), xt_bf16.shape[0], slab, N, vmode,
                                       alpha, mode, _p(out), _p(scratch), _stream()))


def fn_tc_scratch_floats(model, max_rows, M):
    return (int(lib().psvi_fn_tc_scratch_bytes(C.byref(model), max_rows, M)) + 3) // 4


def fn_predictive_tc(model, noise, mu, rho, u, z, v, xt_bf16, yt, slab, N, vmode, alpha, mode, out, scratch):
    """Tensor-core predictive pass for fn with one hidden layer (large regime); same contract as lr_predictive_tc."""
    M = 0 if u is None else u.shape[0]
    _count(10 if mode == 0 else 5)
    _check(lib().psvi_fn_predictive_tc(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v), M,
                                       _p(xt_bf16, torch.bfloat16), _p(yt, torch.int32), xt_bf16.shape[0], slab, N, vmode,
                                       alpha, mode, _p(out), _p(scratch), _stream()))


def fn_nll_tc(model, noise, mu, rho, x_bf16, labels, row_weights, slab, wsum_out, nkl_out, nll_out, scratch):
    """Per-sample weighted NLL sums (and optionally nkl / per-row NLL) of fn on the tensor path."""
    _count(5)
    _check(lib().psvi_fn_nll_tc(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(x_bf16, torch.bfloat16),
                                _p(labels, torch.int32), _p(row_weights), x_bf16.shape[0], slab, _p(wsum_out), _p(nkl_out),
                                _p(nll_out), _p(scratch), _stream()))


def predictive_tc_slabs(model, noise, mu, rho, u, z, v, xt_bf16, yt, batch, first_slab, N, vmode, alpha, mode, out, scratch):
    """PSVI.evaluate over all test batches of this rank in one native call (tcgen05 predictive kernels, one noise slab each)."""
    n_slabs = -(-xt_bf16.shape[0] // int(batch))
    _count((((4 if mode == 0 else 3) if model.n_layers == 1 else (10 if mode == 0 else 5)) + 1) * n_slabs)
    _check(lib().psvi_predictive_tc_slabs(C.byref(model), C.byref(noise), _p(mu), _p(rho), _p(u), _p(z, torch.int32), _p(v),
                                          0 if u is None else u.shape[0], _p(xt_bf16, torch.bfloat16), _p(yt, torch.int32),
                                          xt_bf16.shape[0], int(batch), int(first_slab), N, vmode, alpha, mode, _p(out),
                                          _p(scratch), _stream()))


def fn_data_grad_scratch_floats(model, n_rows):
    return (int(lib().psvi_fn_data_grad_tc_scratch_bytes(C.byref(model), int(n_rows))) + 3) // 4


def fn_data_grad_tc(model, theta, x_bf16, labels, coef, dsum_out, tbar_out, scratch):
    """Data-term gradient of the outer objective over a shard of rows on the tensor path: dsum [S] = sum_r nll[s, r],
    tbar [S][P] = coef[s] * sum_r d nll[s, r] / d theta_s (two tcgen05 passes; include/psvi_b200.h)."""
    _count(6)
    _check(lib().psvi_fn_data_grad_tc(C.byref(model), _p(theta), _p(x_bf16, torch.bfloat16), _p(labels, torch.int32),
                                      x_bf16.shape[0], _p(coef), _p(dsum_out), _p(tbar_out), _p(scratch), _stream()))


def f32_to_bf16(src, dst):
    _count(1)
    _check(lib().psvi_f32_to_bf16(_p(src), _p(dst, torch.bfloat16), src.numel(), _stream()))


def net_pass(model, theta, thetad, x, y, cw, nll=None, tbar=None, tdbar=None, xbar=None, acbar=None, logits=None):
    _count(1)
    _check(lib().psvi_net_pass(C.byref(model), _p(theta), _p(thetad), _p(x), _p(y, torch.int32), _p(cw), x.shape[0],
                               _p(nll), _p(tbar), _p(tdbar), _p(xbar), _p(acbar), _p(logits), _stream()))


def net_pass_gaussian(model, theta, thetad, x, y, cw, tau, nll=None, tbar=None, tdbar=None, xbar=None, acbar=None, ybar=None,
                      outputs=None):
    """Per-sample pass with the Gaussian likelihood of the regressors (y: float targets [R]; one network output)."""
    _count(1)
    _check(lib().psvi_net_pass_gaussian(C.byref(model), _p(theta), _p(thetad), _p(x), _p(y), _p(cw), x.shape[0], float(tau),
                                        _p(nll), _p(tbar), _p(tdbar), _p(xbar), _p(acbar), _p(ybar), _p(outputs), _stream()))


def net_pass_bernoulli(model, theta, thetad, x, y, cw, nll=None, tbar=None, tdbar=None, xbar=None, acbar=None, outputs=None):
    """Per-sample pass with the Bernoulli likelihood on one logit (y: float labels 0. / 1. [R])."""
    _count(1)
    _check(lib().psvi_net_pass_bernoulli(C.byref(model), _p(theta), _p(thetad), _p(x), _p(y), _p(cw), x.shape[0], _p(nll), _p(tbar),
                                         _p(tdbar), _p(xbar), _p(acbar), _p(outputs), _stream()))


def net_predict(model, theta, log_weights, mode, xt, yt, out):
    _count(2)
    ws = _workspace(model, xt.shape[0], theta.device)
    _check(lib().psvi_net_predict(C.byref(model), _p(theta), _p(log_weights), mode, _p(xt), _p(yt, torch.int32),
                                  xt.shape[0], _p(out), _p(ws), _stream()))


def fc_matvec(n, S, base, dg, off, eps_ptr, ld_eps, out_ptr, ld_out):
    """eps_ptr / out_ptr are raw device addresses (views into [S][P] slabs at a layer offset)."""
    _count(1)
    _check(lib().psvi_fc_matvec(n, S, _p(base), _p(dg), _p(off), eps_ptr, ld_eps, out_ptr, ld_out, _stream()))


def fc_outer(n, S, a_ptr, ld_a, eps_ptr, ld_eps, g_base, g_dg, g_off):
    _count(1)
    _check(lib().psvi_fc_outer(n, S, a_ptr, ld_a, eps_ptr, ld_eps, _p(g_base), _p(g_dg), _p(g_off), _stream()))


def fc_sample(n, S, phi_ptr, phid_ptr, eps_ptr, ld_eps, out_ptr, ld_out):
    """Raw device addresses: the layer's block of phi (and of a direction, or None), eps / out at the layer's offset."""
    _count(1)
    _check(lib().psvi_fc_sample(n, S, phi_ptr, phid_ptr, eps_ptr, ld_eps, out_ptr, ld_out, _stream()))


def fc_reparam_grad(n, S, phi_ptr, a_ptr, ld_a, eps_ptr, ld_eps, kl_coef, nkl_coef, g_ptr):
    _count(1)
    _check(lib().psvi_fc_reparam_grad(n, S, phi_ptr, a_ptr, ld_a, eps_ptr, ld_eps, kl_coef, nkl_coef, g_ptr, _stream()))


def fc_reparam_hvp(n, S, phi_ptr, phid_ptr, at_ptr, atd_ptr, ld_a, eps_ptr, ld_eps, h_ptr):
    _count(1)
    _check(lib().psvi_fc_reparam_hvp(n, S, phi_ptr, phid_ptr, at_ptr, atd_ptr, ld_a, eps_ptr, ld_eps, h_ptr, _stream()))


_lenet_ws = {}


def lenet_num_theta():
    return int(lib().psvi_lenet_num_theta())


def lenet_pass(S, theta, thetad, x, y, cw, nll=None, tbar=None, tdbar=None, xbar=None, acbar=None, logits=None):
    """Per-sample lenet pass (forward / gradient / dual) on sampled weights theta [S][P]; x [R][784]."""
    R = x.shape[0]
    n = (int(lib().psvi_lenet_workspace_bytes(S, R)) + 3) // 4
    key = (theta.device, _stream())
    ws = _lenet_ws.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(n, device=theta.device, dtype=torch.float32)
        _lenet_ws[key] = ws
    _count(6 if tbar is None else (16 if thetad is None else 36))
    _check(lib().psvi_lenet_pass(S, _p(theta), _p(thetad), _p(x), _p(y, torch.int32), _p(cw), R, _p(nll), _p(tbar), _p(tdbar),
                                 _p(xbar), _p(acbar), _p(logits), _p(ws), _stream()))


def logits_predict(logits, log_weights, mode, yt, out, probs_out=None):
    S, R, Cc = logits.shape
    _count(1)
    _check(lib().psvi_logits_predict(_p(logits), _p(log_weights), mode, _p(yt, torch.int32), S, R, Cc, _p(out), _p(probs_out),
                                     _stream()))


_fnl_ws = {}


PREC_BF16, PREC_TF32X3, PREC_BF16X3 = 0, 1, 2


def fnl_pass(model, precision, theta, thetad, x, y, cw, nll=None, tbar=None, tdbar=None, xbar=None, acbar=None, logits=None):
    """Large-regime fn pass (batched tcgen05 GEMMs; bf16, tf32x3 or bf16x3 arithmetic) on sampled weights theta [S][P]."""
    R = x.shape[0]
    n = (int(lib().psvi_fnl_workspace_bytes(C.byref(model), R, precision)) + 3) // 4
    key = (theta.device, _stream())
    ws = _fnl_ws.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(n, device=theta.device, dtype=torch.float32)
        _fnl_ws[key] = ws
    _count(8 if tbar is None else (14 if thetad is None else 25))
    _check(lib().psvi_fnl_pass(C.byref(model), precision, _p(theta), _p(thetad), _p(x), _p(y, torch.int32), _p(cw), R, _p(nll),
                               _p(tbar), _p(tdbar), _p(xbar), _p(acbar), _p(logits), _p(ws), _stream()))


# ---- the mean-field family as fused maps over [S][P] slabs (csrc/psvi_family.cu)
def mf_sample(mu, rho, eps):
    S, P = eps.shape
    theta = torch.empty_like(eps)
    _count(1)
    _check(lib().psvi_mf_sample(S, P, _p(mu), _p(rho), _p(eps), _p(theta), _stream()))
    return theta


def mf_tangent(rho, mud, rhod, eps):
    S, P = eps.shape
    thetad = torch.empty_like(eps)
    _count(1)
    _check(lib().psvi_mf_tangent(S, P, _p(rho), _p(mud), _p(rhod), _p(eps), _p(thetad), _stream()))
    return thetad


def mf_reparam_grad(mu, rho, eps, tbar, kl_coef, nkl_coef, mask=None, beta=None, theta=None):
    S, P = eps.shape
    g = torch.empty(2 * P, device=eps.device)
    _count(1)
    _check(lib().psvi_mf_reparam_grad(S, P, _p(mu), _p(rho), _p(eps), _p(tbar), _p(beta), _p(theta), _p(mask), float(kl_coef),
                                      float(nkl_coef), _p(g), _stream()))
    return g


def mf_reparam_hvp(rho, mud, rhod, eps, A_t, A_td, mask=None):
    S, P = eps.shape
    h = torch.empty(2 * P, device=eps.device)
    _count(1)
    _check(lib().psvi_mf_reparam_hvp(S, P, _p(rho), _p(mud), _p(rhod), _p(eps), _p(A_t), _p(A_td), _p(mask), _p(h), _stream()))
    return h


def adam_unroll_step(phi, g, m, v, k, sq2):
    """One step of the unrolled robust Adam -> (phi', m', v') (new tensors)."""
    po, mo, vo = torch.empty_like(phi), torch.empty_like(phi), torch.empty_like(phi)
    _count(1)
    _check(lib().psvi_adam_unroll_step(phi.numel(), k, sq2, _p(phi), _p(g), _p(m), _p(v), _p(po), _p(mo), _p(vo), _stream()))
    return po, mo, vo


def adam_unroll_reverse(pbar, g, m_t, v_t, mbar, vbar, k, sq2):
    """Reverse of that step: updates (mbar, vbar) in place and returns gbar."""
    gbar = torch.empty_like(pbar)
    _count(1)
    _check(lib().psvi_adam_unroll_reverse(pbar.numel(), k, sq2, _p(pbar), _p(g), _p(m_t), _p(v_t), _p(mbar), _p(vbar), _p(gbar),
                                          _stream()))
    return gbar


def mf_nkl_kl(mu, rho, eps, theta, mask=None):
    """-> double tensor [S + 1]: sampled nkl per sample, then the KL."""
    S, P = eps.shape
    out = torch.empty(S + 1, device=eps.device, dtype=torch.float64)
    scratch = torch.empty((int(lib().psvi_mf_nkl_scratch_bytes(S)) + 7) // 8, device=eps.device, dtype=torch.float64)
    _count(2)
    _check(lib().psvi_mf_nkl_kl(S, P, _p(mu), _p(rho), _p(eps), _p(theta), _p(mask), _p(out, torch.float64), _p(scratch, torch.float64),
                                _stream()))
    return out
