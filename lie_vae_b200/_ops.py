"""torch.autograd.Function wrappers over the C ABI (include/lievae.h).

These own allocation, shape / dtype / device checks and raise Python exceptions;
the kernels behind them are hand-written sm_100a CUDA.  CPU tensors are rejected:
this package has no CPU or eager-PyTorch fallback.
"""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _cabi

_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


def _sfx(t):
    try:
        return _SUFFIX[t.dtype]
    except KeyError:
        raise TypeError("lie_vae_b200 kernels take float32 or float64 tensors, got %s" % t.dtype)


def _require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("lie_vae_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                               % t.device.type)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("tensors on different devices: %s vs %s" % (dev, t.device))
    return dev


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """The current CUDA stream of the current device as a ``void*`` for the C ABI (the raw-handle query skips the
    ``torch.cuda.Stream`` object: these wrappers are launch-latency bound at BASELINE configs[1] / configs[2] sizes)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on:
    """``with _on(dev):`` makes ``dev`` the current CUDA device for a launch.  Unlike ``torch.cuda.device`` it costs one
    ``current_device()`` query when ``dev`` already is current (the common case: these wrappers are launch-latency
    bound at BASELINE configs[1] / configs[2] sizes)."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _rows(t, width):
    """(..., width...) -> contiguous (n, width) view/copy and the leading shape."""
    return t.reshape(-1, width).contiguous()


# ------------------------------------------------------------------------------ row ops
def _make_row_op(name, in_shapes, out_shape, bwd_needs_inputs=True):
    """Build a Function for an elementwise map with len(in_shapes) inputs and one output.

    in_shapes / out_shape are the trailing per-sample shapes, e.g. (3,) -> (3, 3).
    Entry points: lv_<name>_fwd_<sfx>(in..., out, n, stream),
                  lv_<name>_bwd_<sfx>([in...,] gout, gin..., n, stream).
    """
    import math
    in_w = [math.prod(s) for s in in_shapes]
    out_w = math.prod(out_shape)

    class _Op(Function):
        @staticmethod
        def forward(ctx, *inputs):
            dev = _require_cuda(*inputs)
            sfx = _sfx(inputs[0])
            for t, s in zip(inputs, in_shapes):
                if t.dtype != inputs[0].dtype:
                    raise TypeError("%s: mixed dtypes" % name)
                if tuple(t.shape[t.dim() - len(s):]) != tuple(s):
                    raise ValueError("%s: expected trailing shape %s, got %s" % (name, tuple(s), tuple(t.shape)))
            lead = inputs[0].shape[:inputs[0].dim() - len(in_shapes[0])]
            flat = [_rows(t, w) for t, w in zip(inputs, in_w)]
            n = flat[0].shape[0]
            for f in flat:
                if f.shape[0] != n:
                    raise ValueError("%s: inputs disagree on the batch size" % name)
            out = torch.empty((n, out_w), dtype=inputs[0].dtype, device=dev)
            with _on(dev):
                _cabi.call("lv_%s_fwd_%s" % (name, sfx), *[_cabi.ptr(f) for f in flat], _cabi.ptr(out), n, _stream())
            ctx.save_for_backward(*(flat if bwd_needs_inputs else []))
            ctx.meta = (sfx, n, lead, [t.shape for t in inputs])
            return out.reshape(*lead, *out_shape)

        @staticmethod
        @once_differentiable
        def backward(ctx, gout):
            sfx, n, lead, shapes = ctx.meta
            flat = list(ctx.saved_tensors)
            g = gout.reshape(-1, out_w).contiguous()
            dev = g.device
            gins = [torch.empty((n, w), dtype=g.dtype, device=dev) for w in in_w]
            with _on(dev):
                _cabi.call("lv_%s_bwd_%s" % (name, sfx), *[_cabi.ptr(f) for f in flat], _cabi.ptr(g),
                           *[_cabi.ptr(x) for x in gins], n, _stream())
            return tuple(x.reshape(s) for x, s in zip(gins, shapes))

    _Op.__name__ = "LV_" + name
    return _Op


Hat = _make_row_op("hat", [(3,)], (3, 3), bwd_needs_inputs=False)
Vee = _make_row_op("vee", [(3, 3)], (3,), bwd_needs_inputs=False)
Rodrigues = _make_row_op("rodrigues", [(3,)], (3, 3))
LogMap = _make_row_op("log_map", [(3, 3)], (3, 3))
QuatToMat = _make_row_op("quat_to_mat", [(4,)], (3, 3))
MatToQuat = _make_row_op("mat_to_quat", [(3, 3)], (4,))
QuatToEazyz = _make_row_op("quat_to_eazyz", [(4,)], (3,))
MatToEazyz = _make_row_op("mat_to_eazyz", [(3, 3)], (3,))
S2S1Rodrigues = _make_row_op("s2s1_rodrigues", [(3,), (2,)], (3, 3))
S2S2GramSchmidt = _make_row_op("s2s2_gram_schmidt", [(3,), (3,)], (3, 3))
VectorToEazyz = _make_row_op("vector_to_eazyz", [(3,)], (3,))


def sum_leading(t):
    """(n, ...) -> (...) summed over the first axis (gradient reduction over samples)."""
    dev = _require_cuda(t)
    n = t.shape[0]
    if n == 1:
        return t[0]
    flat = t.reshape(n, -1).contiguous()
    out = torch.empty(flat.shape[1], dtype=t.dtype, device=dev)
    with _on(dev):
        _cabi.call("lv_sum_leading_%s" % _sfx(t), _cabi.ptr(flat), _cabi.ptr(out), n, flat.shape[1], _stream())
    return out.reshape(t.shape[1:])


class LogSumExpLeading(Function):
    """(n, ...) -> (...): log-sum-exp over the first axis, one pass (running max / sum) per column."""

    @staticmethod
    def forward(ctx, x):
        dev = _require_cuda(x)
        sfx = _sfx(x)
        n = x.shape[0]
        if n == 0:
            raise ValueError("logsumexp over an empty axis")
        flat = x.reshape(n, -1).contiguous()
        out = torch.empty(flat.shape[1], dtype=x.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_logsumexp_leading_fwd_" + sfx, _cabi.ptr(flat), _cabi.ptr(out), n, flat.shape[1], _stream())
        ctx.save_for_backward(flat, out)
        ctx.shape = tuple(x.shape)
        return out.reshape(x.shape[1:])

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        flat, out = ctx.saved_tensors
        n, inner = flat.shape
        g = gout.reshape(-1).contiguous()
        gin = torch.empty_like(flat)
        with _on(flat.device):
            _cabi.call("lv_logsumexp_leading_bwd_" + _sfx(flat), _cabi.ptr(flat), _cabi.ptr(out), _cabi.ptr(g), _cabi.ptr(gin),
                       n, inner, _stream())
        return gin.reshape(ctx.shape)


# ------------------------------------------------------------------------------ fused SO(3) reparameterize
class SO3Reparam(Function):
    """(mu (B,3,3), sigma (B,3), eps (n,B,3), k) -> z (n,B,3,3), log_q (n,B).   float32 (production) or float64."""

    @staticmethod
    def forward(ctx, mu, sigma, eps, k):
        dev = _require_cuda(mu, sigma, eps)
        sfx = _sfx(mu)
        for t in (sigma, eps):
            if t.dtype != mu.dtype:
                raise TypeError("so3_reparameterize: mu, sigma, eps must share a dtype, got %s / %s" % (mu.dtype, t.dtype))
        if mu.dim() != 3 or tuple(mu.shape[1:]) != (3, 3):
            raise ValueError("mu must be (B,3,3), got %s" % (tuple(mu.shape),))
        B = mu.shape[0]
        if tuple(sigma.shape) != (B, 3):
            raise ValueError("sigma must be (B,3), got %s" % (tuple(sigma.shape),))
        if eps.dim() != 3 or tuple(eps.shape[1:]) != (B, 3):
            raise ValueError("eps must be (n,B,3), got %s" % (tuple(eps.shape),))
        n = eps.shape[0]
        mu_c, sg_c, ep_c = mu.contiguous(), sigma.contiguous(), eps.contiguous()
        z = torch.empty((n, B, 3, 3), dtype=mu.dtype, device=dev)
        log_q = torch.empty((n, B), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_fwd_" + sfx, _cabi.ptr(mu_c), _cabi.ptr(sg_c), _cabi.ptr(ep_c), _cabi.ptr(z),
                       _cabi.ptr(log_q), n, B, int(k), _stream())
        ctx.save_for_backward(mu_c, sg_c, ep_c)
        ctx.k = int(k)
        ctx.set_materialize_grads(False)       # unused outputs arrive as None, not as zero tensors
        return z, log_q

    @staticmethod
    @once_differentiable
    def backward(ctx, gz, glq):
        mu, sigma, eps = ctx.saved_tensors
        n, B = eps.shape[0], eps.shape[1]
        dev = mu.device
        gz = None if gz is None else gz.contiguous()
        glq = None if glq is None else glq.contiguous()
        gmu = torch.empty((n, B, 3, 3), dtype=mu.dtype, device=dev)
        gsg = torch.empty((n, B, 3), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_bwd_" + _sfx(mu), _cabi.ptr(mu), _cabi.ptr(sigma), _cabi.ptr(eps), _cabi.ptr(gz),
                       _cabi.ptr(glq), _cabi.ptr(gmu), _cabi.ptr(gsg), n, B, ctx.k, _stream())
        return sum_leading(gmu), sum_leading(gsg), None, None


class SO3ReparamEazyz(Function):
    """(mu, sigma, eps, k) -> angles (n,B,3), log_q (n,B): reparameterize fused with matrix -> ZYZ Euler.

    The pose z = mu @ exp(hat(eps*sigma)) stays in registers; only its Euler angles (what VAE.decode hands
    to the action decoder, ``experiments/vae.py:182``) and the log-density are written.   float32.
    """

    @staticmethod
    def forward(ctx, mu, sigma, eps, k):
        dev = _require_cuda(mu, sigma, eps)
        sfx = _sfx(mu)
        for t in (sigma, eps):
            if t.dtype != mu.dtype:
                raise TypeError("so3_reparameterize_eazyz: mu, sigma, eps must share a dtype, got %s / %s" % (mu.dtype, t.dtype))
        if mu.dim() != 3 or tuple(mu.shape[1:]) != (3, 3):
            raise ValueError("mu must be (B,3,3), got %s" % (tuple(mu.shape),))
        B = mu.shape[0]
        if tuple(sigma.shape) != (B, 3):
            raise ValueError("sigma must be (B,3), got %s" % (tuple(sigma.shape),))
        if eps.dim() != 3 or tuple(eps.shape[1:]) != (B, 3):
            raise ValueError("eps must be (n,B,3), got %s" % (tuple(eps.shape),))
        n = eps.shape[0]
        mu_c, sg_c, ep_c = mu.contiguous(), sigma.contiguous(), eps.contiguous()
        angles = torch.empty((n, B, 3), dtype=mu.dtype, device=dev)
        log_q = torch.empty((n, B), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_eazyz_fwd_" + sfx, _cabi.ptr(mu_c), _cabi.ptr(sg_c), _cabi.ptr(ep_c), None,
                       _cabi.ptr(angles), _cabi.ptr(log_q), n, B, int(k), _stream())
        ctx.save_for_backward(mu_c, sg_c, ep_c)
        ctx.k = int(k)
        ctx.set_materialize_grads(False)       # unused outputs arrive as None, not as zero tensors
        return angles, log_q

    @staticmethod
    @once_differentiable
    def backward(ctx, gang, glq):
        mu, sigma, eps = ctx.saved_tensors
        n, B = eps.shape[0], eps.shape[1]
        dev = mu.device
        gang = torch.zeros((n, B, 3), dtype=mu.dtype, device=dev) if gang is None else gang.contiguous()
        glq = None if glq is None else glq.contiguous()
        gmu = torch.empty((n, B, 3, 3), dtype=mu.dtype, device=dev)
        gsg = torch.empty((n, B, 3), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_eazyz_bwd_" + _sfx(mu), _cabi.ptr(mu), _cabi.ptr(sigma), _cabi.ptr(eps), None, _cabi.ptr(gang),
                       _cabi.ptr(glq), _cabi.ptr(gmu), _cabi.ptr(gsg), n, B, ctx.k, _stream())
        return sum_leading(gmu), sum_leading(gsg), None, None


class SO3ReparamPhilox(Function):
    """(mu (B,3,3), sigma (B,3), n, k, seed, offset, euler) -> (pose, log_q) with IN-KERNEL noise.

    eps ~ N(0,1) of the reference (``reparameterize.py:137-141``) is generated inside the kernel by Philox4x32-10
    keyed by ``seed`` at counter ``offset + flat sample index`` and regenerated by the backward: it is never stored.
    ``pose`` is z (n,B,3,3), or its ZYZ Euler angles (n,B,3) with ``euler``.  ``philox_normal`` returns the same eps.
    """

    @staticmethod
    def forward(ctx, mu, sigma, n, k, seed, offset, euler):
        dev = _require_cuda(mu, sigma)
        sfx = _sfx(mu)
        if sigma.dtype != mu.dtype:
            raise TypeError("so3_reparameterize: mu, sigma must share a dtype, got %s / %s" % (mu.dtype, sigma.dtype))
        if mu.dim() != 3 or tuple(mu.shape[1:]) != (3, 3):
            raise ValueError("mu must be (B,3,3), got %s" % (tuple(mu.shape),))
        B = mu.shape[0]
        if tuple(sigma.shape) != (B, 3):
            raise ValueError("sigma must be (B,3), got %s" % (tuple(sigma.shape),))
        n = int(n)
        mu_c, sg_c = mu.contiguous(), sigma.contiguous()
        pose = torch.empty((n, B, 3) if euler else (n, B, 3, 3), dtype=mu.dtype, device=dev)
        log_q = torch.empty((n, B), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_philox_fwd_" + sfx, _cabi.ptr(mu_c), _cabi.ptr(sg_c), int(seed), int(offset),
                       None if euler else _cabi.ptr(pose), _cabi.ptr(pose) if euler else None, _cabi.ptr(log_q), n, B, int(k), _stream())
        ctx.save_for_backward(mu_c, sg_c)
        ctx.meta = (n, int(k), int(seed), int(offset), bool(euler))
        ctx.set_materialize_grads(False)       # unused outputs arrive as None, not as zero tensors
        return pose, log_q

    @staticmethod
    @once_differentiable
    def backward(ctx, gpose, glq):
        mu, sigma = ctx.saved_tensors
        n, k, seed, offset, euler = ctx.meta
        B, dev = mu.shape[0], mu.device
        if euler and gpose is None:
            gpose = torch.zeros((n, B, 3), dtype=mu.dtype, device=dev)
        gpose = None if gpose is None else gpose.contiguous()
        glq = None if glq is None else glq.contiguous()
        gmu = torch.empty((n, B, 3, 3), dtype=mu.dtype, device=dev)
        gsg = torch.empty((n, B, 3), dtype=mu.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_so3_reparam_philox_bwd_" + _sfx(mu), _cabi.ptr(mu), _cabi.ptr(sigma), seed, offset,
                       None if euler else _cabi.ptr(gpose), _cabi.ptr(gpose) if euler else None, _cabi.ptr(glq),
                       _cabi.ptr(gmu), _cabi.ptr(gsg), n, B, k, _stream())
        return sum_leading(gmu), sum_leading(gsg), None, None, None, None, None


def philox_normal(rows, seed, offset=0, dtype=torch.float32, device="cuda"):
    """(rows, 3) standard normals: exactly the eps SO3ReparamPhilox uses for flat samples offset .. offset + rows - 1."""
    out = torch.empty((int(rows), 3), dtype=dtype, device=device)
    dev = _require_cuda(out)
    with _on(dev):
        _cabi.call("lv_philox_normal_" + _sfx(out), _cabi.ptr(out), int(rows), int(seed), int(offset), _stream())
    return out


# ------------------------------------------------------------------------------ C++ autograd fast path (csrc/torch_binding.cpp)
_TORCH_EXT = False        # False: not looked for yet; None: absent / disabled; else the extension module


def torch_ext():
    """The in-tree C++ autograd extension (``python -m lie_vae_b200._build``), or None.  It wraps the same C-ABI entry points
    as the Python Functions of this module, without re-entering Python in the backward (~115 us less host time per
    forward + backward pair); ``LIEVAE_NO_TORCH_EXT=1`` keeps the Python Functions."""
    global _TORCH_EXT
    if _TORCH_EXT is False:
        _TORCH_EXT = None
        import os
        from . import _build
        if os.environ.get("LIEVAE_NO_TORCH_EXT") != "1" and os.path.exists(_build.TORCH_EXT_LIB) and not _build.torch_ext_is_stale():
            import importlib.util
            _cabi.lib()                   # the kernels' library first: the extension links against it by soname
            spec = importlib.util.spec_from_file_location(_build.TORCH_EXT_NAME, _build.TORCH_EXT_LIB)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            if mod.abi_version() != _cabi.lib().lv_version():
                raise RuntimeError("lie_vae_b200: %s was built against another liblievae_sm100a.so; rebuild with "
                                   "`python -m lie_vae_b200._build --force`" % _build.TORCH_EXT_LIB)
            _TORCH_EXT = mod
    return _TORCH_EXT


def _fast(*tensors):
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors) and torch_ext() is not None


def so3_reparam(mu, sigma, eps, k, euler=False):
    """(pose, log_q) of the fused reparameterize kernel; ``pose`` = z (n,B,3,3), or its ZYZ Euler angles (n,B,3) with ``euler``."""
    if _fast(mu, sigma, eps):
        # same argument errors as the Python Functions below (the extension re-checks and would raise RuntimeError)
        if mu.dim() != 3 or tuple(mu.shape[1:]) != (3, 3):
            raise ValueError("mu must be (B,3,3), got %s" % (tuple(mu.shape),))
        if tuple(sigma.shape) != (mu.shape[0], 3):
            raise ValueError("sigma must be (B,3), got %s" % (tuple(sigma.shape),))
        if eps.dim() != 3 or tuple(eps.shape[1:]) != (mu.shape[0], 3):
            raise ValueError("eps must be (n,B,3), got %s" % (tuple(eps.shape),))
        pose, log_q = torch_ext().so3_reparam(mu, sigma, eps, int(k), bool(euler))
        return pose, log_q
    return (SO3ReparamEazyz if euler else SO3Reparam).apply(mu, sigma, eps, k)


HEAD_MODES = {"alg": 0, "q": 1, "s2s2": 2, "s2s1": 3}      # mean maps the fused head kernels know (mean-head rows: 3, 4, 6, 5)
HEAD_MAX_DIN = 32


class SO3HeadReparam(Function):
    """(h (B,Din), Wm (Dm,Din), bm (Dm), Ws (3,Din), bs (3), eps (n,B,3), mode, k, euler) -> (pose, log_q, mu, sigma).

    Encoder heads fused into the reparameterize kernel: mu = mean_map(Wm h + bm), sigma = softplus(Ws h + bs), then the
    sampler.  ``pose`` is z (n,B,3,3), or its ZYZ Euler angles (n,B,3) with ``euler``.  mu (B,3,3) and sigma (B,3)
    (the modules' ``mu_lie`` / ``sigma`` attributes) are differentiable outputs like the other two: gradients that reach
    them from outside the sampler are folded into the same backward kernel.   float32.
    """

    @staticmethod
    def forward(ctx, h, Wm, bm, Ws, bs, eps, mode, k, euler):
        dev = _require_cuda(h, Wm, bm, Ws, bs, eps)
        for t in (h, Wm, bm, Ws, bs, eps):
            if t.dtype != torch.float32:
                raise TypeError("so3_head_reparameterize is float32 only, got %s" % t.dtype)
        m = HEAD_MODES[mode]
        dm = (3, 4, 6, 5)[m]
        if h.dim() != 2 or h.shape[1] > HEAD_MAX_DIN:
            raise ValueError("h must be (B, Din <= %d), got %s" % (HEAD_MAX_DIN, tuple(h.shape)))
        B, Din = h.shape
        if tuple(Wm.shape) != (dm, Din) or tuple(bm.shape) != (dm,) or tuple(Ws.shape) != (3, Din) or tuple(bs.shape) != (3,):
            raise ValueError("head parameters must be (%d,%d), (%d,), (3,%d), (3,); got %s %s %s %s"
                             % (dm, Din, dm, Din, tuple(Wm.shape), tuple(bm.shape), tuple(Ws.shape), tuple(bs.shape)))
        if eps.dim() != 3 or tuple(eps.shape[1:]) != (B, 3):
            raise ValueError("eps must be (n,B,3), got %s" % (tuple(eps.shape),))
        n = eps.shape[0]
        saved = tuple(t.contiguous() for t in (h, Wm, bm, Ws, bs, eps))
        f32 = dict(dtype=torch.float32, device=dev)
        mu, sigma = torch.empty((B, 3, 3), **f32), torch.empty((B, 3), **f32)
        pose = torch.empty((n, B, 3) if euler else (n, B, 3, 3), **f32)
        log_q = torch.empty((n, B), **f32)
        with _on(dev):
            _cabi.call("lv_so3_head_reparam_fwd_f32", *[_cabi.ptr(t) for t in saved], _cabi.ptr(mu), _cabi.ptr(sigma),
                       None if euler else _cabi.ptr(pose), _cabi.ptr(pose) if euler else None, _cabi.ptr(log_q),
                       n, B, Din, m, int(k), _stream())
        ctx.save_for_backward(*saved)
        ctx.meta = (m, int(k), bool(euler), dm)
        ctx.set_materialize_grads(False)       # unused outputs arrive as None, not as zero tensors
        return pose, log_q, mu, sigma

    @staticmethod
    @once_differentiable
    def backward(ctx, gpose, glq, gmu, gsigma):
        saved = ctx.saved_tensors
        h, eps = saved[0], saved[5]
        m, k, euler, dm = ctx.meta
        n, B, Din = eps.shape[0], h.shape[0], h.shape[1]
        dev = h.device
        f32 = dict(dtype=torch.float32, device=dev)
        if gpose is None and glq is None and gmu is None and gsigma is None:
            return tuple(torch.zeros_like(t) for t in saved[:5]) + (None, None, None, None)
        gmu = None if gmu is None else gmu.contiguous()
        gsigma = None if gsigma is None else gsigma.contiguous()
        if euler and gpose is None:
            gpose = torch.zeros((n, B, 3), **f32)
        gpose = None if gpose is None else gpose.contiguous()
        glq = None if glq is None else glq.contiguous()
        gh = torch.empty((n, B, Din), **f32)
        gwb = torch.empty((dm + 3, Din + 1), **f32)
        with _on(dev):
            nws = _cabi.lib().lv_so3_head_reparam_bwd_workspace_floats(n, B, Din, m)
            ws = torch.empty(max(nws, 1), **f32)
            _cabi.call("lv_so3_head_reparam_bwd_f32", *[_cabi.ptr(t) for t in saved],
                       None if euler else _cabi.ptr(gpose), _cabi.ptr(gpose) if euler else None, _cabi.ptr(glq),
                       _cabi.ptr(gmu), _cabi.ptr(gsigma), _cabi.ptr(gh), _cabi.ptr(gwb), _cabi.ptr(ws), nws, n, B, Din, m, k, _stream())
        return sum_leading(gh), gwb[:dm, :Din], gwb[:dm, Din], gwb[dm:, :Din], gwb[dm:, Din], None, None, None, None


# ------------------------------------------------------------------------------ Wigner-D action
class WignerApply(Function):
    """angles (N,3), spectrum ((M,C) shared | (N,M,C)) -> (N,M,C), degrees lmin..lmax.   float32."""

    @staticmethod
    def forward(ctx, angles, spectrum, lmin, lmax, transpose):
        dev = _require_cuda(angles, spectrum)
        if angles.dtype != torch.float32 or spectrum.dtype != torch.float32:
            raise TypeError("wigner apply is float32 only, got %s / %s" % (angles.dtype, spectrum.dtype))
        if angles.dim() != 2 or angles.shape[1] != 3:
            raise ValueError("angles must be (N,3)")
        N = angles.shape[0]
        M = (lmax + 1) ** 2 - lmin ** 2
        shared = spectrum.dim() == 2
        if shared:
            if spectrum.shape[0] != M:
                raise ValueError("spectrum must have %d rows for degrees %d..%d, got %s" % (M, lmin, lmax, tuple(spectrum.shape)))
        elif spectrum.dim() != 3 or spectrum.shape[0] != N or spectrum.shape[1] != M:
            raise ValueError("spectrum must be (N=%d, M=%d, C), got %s" % (N, M, tuple(spectrum.shape)))
        C = spectrum.shape[-1]
        a_c, s_c = angles.contiguous(), spectrum.contiguous()
        out = torch.empty((N, M, C), dtype=torch.float32, device=dev)
        with _on(dev):
            _cabi.call("lv_wigner_apply_fwd_f32", _cabi.ptr(a_c), _cabi.ptr(s_c), _cabi.ptr(out), N, lmin, lmax, C,
                       int(shared), int(bool(transpose)), _stream())
        ctx.save_for_backward(a_c, s_c)
        ctx.meta = (N, lmin, lmax, C, shared, bool(transpose))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        a_c, s_c = ctx.saved_tensors
        N, lmin, lmax, C, shared, transpose = ctx.meta
        dev = a_c.device
        g = gout.contiguous()
        gang = torch.empty((N, 3), dtype=torch.float32, device=dev)
        gspec = torch.empty_like(s_c)
        with _on(dev):
            ws, nws = None, 0
            if shared:
                nws = _cabi.lib().lv_wigner_bwd_workspace_floats(N, lmin, lmax, C)
                if nws < 0:
                    raise RuntimeError("lv_wigner_bwd_workspace_floats: " + _cabi.last_error())
                ws = torch.empty(max(nws, 1), dtype=torch.float32, device=dev)
            _cabi.call("lv_wigner_apply_bwd_f32", _cabi.ptr(a_c), _cabi.ptr(s_c), _cabi.ptr(g), _cabi.ptr(gang),
                       _cabi.ptr(gspec), _cabi.ptr(ws), nws, N, lmin, lmax, C, int(shared), int(transpose), _stream())
        return gang, gspec, None, None, None


# ------------------------------------------------------------------------------ generic Wigner action (any degree, f32 / f64)
_J_TABLES = {}


def _j_table(lmax, dtype, device):
    """Packed dense J_0 | ... | J_lmax on `device` (cached; the C ABI takes it as a caller-owned pointer)."""
    key = (int(lmax), dtype, str(device))
    t = _J_TABLES.get(key)
    if t is None:
        import numpy as np
        from .jmatrix import j_table
        t = torch.from_numpy(j_table(int(lmax), np.float64)).to(device=device, dtype=dtype)
        _J_TABLES[key] = t
    return t


def generic_max_degree():
    return int(_cabi.lib().lv_wigner_generic_max_degree())


class WignerApplyGeneric(Function):
    """Same contract as WignerApply for any degree range up to generic_max_degree() and float32 / float64."""

    @staticmethod
    def forward(ctx, angles, spectrum, lmin, lmax, transpose):
        dev = _require_cuda(angles, spectrum)
        if angles.dtype != spectrum.dtype:
            raise TypeError("angles and spectrum must share a dtype, got %s / %s" % (angles.dtype, spectrum.dtype))
        sfx = _sfx(angles)
        if angles.dim() != 2 or angles.shape[1] != 3:
            raise ValueError("angles must be (N,3)")
        if lmax > generic_max_degree():
            raise NotImplementedError("degree %d > %d is not supported" % (lmax, generic_max_degree()))
        N = angles.shape[0]
        M = (lmax + 1) ** 2 - lmin ** 2
        shared = spectrum.dim() == 2
        if shared:
            if spectrum.shape[0] != M:
                raise ValueError("spectrum must have %d rows for degrees %d..%d, got %s" % (M, lmin, lmax, tuple(spectrum.shape)))
        elif spectrum.dim() != 3 or spectrum.shape[0] != N or spectrum.shape[1] != M:
            raise ValueError("spectrum must be (N=%d, M=%d, C), got %s" % (N, M, tuple(spectrum.shape)))
        C = spectrum.shape[-1]
        a_c, s_c = angles.contiguous(), spectrum.contiguous()
        jt = _j_table(lmax, angles.dtype, dev)
        out = torch.empty((N, M, C), dtype=angles.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_wigner_generic_fwd_" + sfx, _cabi.ptr(a_c), _cabi.ptr(s_c), _cabi.ptr(jt), _cabi.ptr(out), N, lmin, lmax,
                       C, int(shared), int(bool(transpose)), _stream())
        ctx.save_for_backward(a_c, s_c, jt)
        ctx.meta = (N, M, lmin, lmax, C, shared, bool(transpose), sfx)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        a_c, s_c, jt = ctx.saved_tensors
        N, M, lmin, lmax, C, shared, transpose, sfx = ctx.meta
        dev = a_c.device
        g = gout.contiguous()
        parts = torch.empty((N, C, 3), dtype=a_c.dtype, device=dev)
        gspec = torch.empty((N, M, C), dtype=a_c.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_wigner_generic_bwd_" + sfx, _cabi.ptr(a_c), _cabi.ptr(s_c), _cabi.ptr(jt), _cabi.ptr(g), _cabi.ptr(parts),
                       _cabi.ptr(gspec), N, lmin, lmax, C, int(shared), int(transpose), _stream())
        return parts.sum(1), (gspec.sum(0) if shared else gspec), None, None, None


def wigner_recon_sse(angles, item_rep, x, lmax, transpose=False):
    """(N) sums of squares between the action output D(angles_i) item_rep and x[i % B] (x: (B, M, C)) -- the reconstruction
    term of ``VAE.log_likelihood`` for the toy deconv, with y never written.  No gradient (evaluation path)."""
    dev = _require_cuda(angles, item_rep, x)
    for t in (angles, item_rep, x):
        if t.dtype != torch.float32:
            raise TypeError("wigner_recon_sse is float32 only, got %s" % t.dtype)
    M, C = item_rep.shape
    if angles.dim() != 2 or angles.shape[1] != 3 or M != (lmax + 1) ** 2 or x.dim() != 3 or tuple(x.shape[1:]) != (M, C):
        raise ValueError("angles (N,3), item_rep ((lmax+1)^2, C) and x (B, M, C) expected")
    N, B = angles.shape[0], x.shape[0]
    out = torch.empty(N, dtype=torch.float32, device=dev)
    with _on(dev):
        _cabi.call("lv_wigner_recon_sse_f32", _cabi.ptr(angles.detach().contiguous()), _cabi.ptr(item_rep.detach().contiguous()),
                   _cabi.ptr(x.detach().contiguous()), _cabi.ptr(out), N, B, 0, int(lmax), C, int(bool(transpose)), _stream())
    return out


class EquivarianceSqDist(Function):
    """theta (n), R (n,3,3), R2 (n,3,3) -> (n) squared Frobenius distances || Rx(theta) R - R2 ||^2, Rx the rotation about the x
    axis built as ``s2s1rodrigues(e_x, (cos, sin))`` -- the SO(3) part of ``EquivarianceLoss.forward``
    (``losses/equivariance_loss.py:27-36``) in one kernel per direction; gradients w.r.t. R and R2."""

    @staticmethod
    def forward(ctx, theta, R, R2):
        dev = _require_cuda(theta, R, R2)
        sfx = _sfx(R)
        if theta.dtype != R.dtype or R2.dtype != R.dtype:
            raise TypeError("equivariance_sqdist: mixed dtypes")
        if tuple(R.shape[-2:]) != (3, 3) or R2.shape != R.shape or theta.shape != R.shape[:-2]:
            raise ValueError("equivariance_sqdist: theta (n), R (n,3,3), R2 (n,3,3) expected, got %s %s %s"
                             % (tuple(theta.shape), tuple(R.shape), tuple(R2.shape)))
        th, a, b = theta.reshape(-1).contiguous(), _rows(R, 9), _rows(R2, 9)
        n = th.shape[0]
        diff = torch.empty(n, dtype=R.dtype, device=dev)
        resid = torch.empty((n, 9), dtype=R.dtype, device=dev)
        with _on(dev):
            _cabi.call("lv_equivariance_sqdist_fwd_%s" % sfx, _cabi.ptr(th), _cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(diff), _cabi.ptr(resid), n, _stream())
        ctx.save_for_backward(th, resid)
        ctx.meta = (sfx, n, R.shape)
        ctx.set_materialize_grads(False)
        return diff.reshape(theta.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, gdiff):
        if gdiff is None:
            return None, None, None
        sfx, n, shape = ctx.meta
        th, resid = ctx.saved_tensors
        g = gdiff.reshape(-1).contiguous()
        gR, gR2 = torch.empty_like(resid), torch.empty_like(resid)
        with _on(g.device):
            _cabi.call("lv_equivariance_sqdist_bwd_%s" % sfx, _cabi.ptr(th), _cabi.ptr(resid), _cabi.ptr(g), _cabi.ptr(gR), _cabi.ptr(gR2), n, _stream())
        return None, gR.reshape(shape), gR2.reshape(shape)


FAST_MAX_DEGREE = 8   # degrees covered by the unrolled, packed float32 kernels


def wigner_apply(angles, spectrum, lmin, lmax, transpose=False):
    """Dispatch: unrolled packed kernels for float32 and degrees <= 8 (through the C++ autograd binding when it is built),
    the generic kernels otherwise."""
    if angles.dtype == torch.float32 and spectrum.dtype == torch.float32 and lmax <= FAST_MAX_DEGREE:
        if _fast(angles, spectrum) and angles.dim() == 2 and angles.shape[1] == 3:
            M = (lmax + 1) ** 2 - lmin ** 2
            if spectrum.dim() == 2 and spectrum.shape[0] == M or spectrum.dim() == 3 and tuple(spectrum.shape[:2]) == (angles.shape[0], M):
                return torch_ext().wigner_apply(angles, spectrum, int(lmin), int(lmax), bool(transpose))
            # wrong shapes: the Python Function below raises the documented ValueError
        return WignerApply.apply(angles, spectrum, lmin, lmax, transpose)
    return WignerApplyGeneric.apply(angles, spectrum, lmin, lmax, transpose)


# ------------------------------------------------------------------------------ the action's consumer on the tensor cores
def round_tf32(t):
    """Nearest-TF32 copy of a float32 tensor (cvt.rna: what cuBLAS / cuDNN apply to TF32 operands)."""
    dev = _require_cuda(t)
    src = t.contiguous()
    out = torch.empty_like(src)
    with _on(dev):
        _cabi.call("lv_round_tf32_f32", _cabi.ptr(src), _cabi.ptr(out), src.numel(), _stream())
    return out


def gemm_tf32(a, bt, bias=None, bias_div=1, out=None):
    """out (M,N) = a (M,K) @ bt (N,K)^T + bias[col // bias_div]: tcgen05 TF32 tensor-core GEMM (lv_gemm_tf32_f32).
    ``a`` / ``bt`` / ``out`` may be row-strided views (last dim contiguous).  ``a`` is rounded to TF32 in the kernel; pass
    ``bt`` through ``round_tf32`` for unbiased products (its low 13 mantissa bits are otherwise ignored)."""
    dev = _require_cuda(a, bt, bias)
    for t in (a, bt) + ((bias,) if bias is not None else ()):
        if t.dtype != torch.float32:
            raise TypeError("gemm_tf32 is float32 (TF32 tensor-core arithmetic), got %s" % t.dtype)
    if a.dim() != 2 or bt.dim() != 2 or a.shape[1] != bt.shape[1] or a.stride(1) != 1 or bt.stride(1) != 1:
        raise ValueError("gemm_tf32: a (M,K) and bt (N,K) with contiguous rows expected, got %s %s" % (tuple(a.shape), tuple(bt.shape)))
    M, K = a.shape
    N = bt.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (M, N) or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("gemm_tf32: out must be a float32 (M,N) view with contiguous rows")
    with _on(dev):
        _cabi.call("lv_gemm_tf32_f32", _cabi.ptr(a), a.stride(0), _cabi.ptr(bt), bt.stride(0), _cabi.ptr(bias), int(bias_div),
                   _cabi.ptr(out), out.stride(0), M, N, K, _stream())
    return out


ACTION_GEMM_CHUNK = 8192      # samples per chunk: y of a chunk (3 240 B/sample for l <= 8, C = 10) stays L2-resident
# The weight gradient y^T g_out of the fused op is a cuBLAS call (its operands are MN-major for the tcgen05 kernel above).  True: TF32
# operands there too -- what cuDNN does for the reference's ConvTranspose2d weight gradient under PyTorch's defaults, and what
# the forward and the data gradient of this op use; False: FP32 (twice the time: 64 vs 133 TFLOP/s).
ACTION_GEMM_WGRAD_TF32 = True


class ActionGemm(Function):
    """angles (N,3), item_rep (M,C), weight (M*C, Nout), bias -> (N, Nout) = wigner_apply(angles, item_rep).view(N, M*C) @ weight + bias.

    The Wigner action fused with its consumer: the first ConvTranspose2d of DeconvNet on the 1x1 input
    (``experiments/nets.py:65-66``; weight viewed (M*C, 16*hidden), bias_div = 16) or the first Linear of ActionNet's MLP
    (``decoders.py:39-41``; weight = linear.weight.t(), bias_div = 1).  Samples are processed in chunks whose action output y
    stays in L2: per chunk the Wigner forward kernel writes y into a reused buffer and the tcgen05 TF32 GEMM reads it back
    from L2; y never round-trips through HBM and is not kept for the backward, which recomputes it per chunk, runs the
    data gradient g_y = g_out W^T on the same tcgen05 kernel (TF32, like cuDNN's default for the reference's convolution),
    feeds it from L2 straight into the Wigner backward kernel, and leaves the weight gradient y^T g_out to cuBLAS (TF32 operands by default, ACTION_GEMM_WGRAD_TF32).
    """

    @staticmethod
    def forward(ctx, angles, item_rep, weight, bias, bias_div, lmax, transpose, chunk):
        dev = _require_cuda(angles, item_rep, weight, bias)
        if angles.dim() != 2 or angles.shape[1] != 3 or item_rep.dim() != 2:
            raise ValueError("angles (N,3) and a shared item_rep (M,C) expected")
        N, (Mh, C) = angles.shape[0], item_rep.shape
        if Mh != (lmax + 1) ** 2 or tuple(weight.shape[:1]) != (Mh * C,) or weight.dim() != 2:
            raise ValueError("weight must be (M*C = %d, Nout), got %s" % (Mh * C, tuple(weight.shape)))
        K, Nout = weight.shape
        a_c, s_c = angles.contiguous(), item_rep.contiguous()
        ldb = (K + 3) // 4 * 4
        wt = torch.zeros((Nout, ldb), dtype=torch.float32, device=dev)       # K-major weight, rows padded to 16 bytes for TMA
        wt[:, :K] = weight.t()
        with _on(dev):
            _cabi.call("lv_round_tf32_f32", _cabi.ptr(wt), _cabi.ptr(wt), wt.numel(), _stream())     # nearest TF32, once per call
        out = torch.empty((N, Nout), dtype=torch.float32, device=dev)
        chunk = max(1, min(int(chunk), N)) if N else 1
        ybuf = torch.empty((chunk, K), dtype=torch.float32, device=dev)
        with _on(dev):
            st = _stream()
            for lo in range(0, N, chunk):
                n = min(chunk, N - lo)
                _cabi.call("lv_wigner_apply_fwd_f32", _cabi.ptr(a_c[lo:lo + n]), _cabi.ptr(s_c), _cabi.ptr(ybuf), n, 0, lmax, C, 1, int(bool(transpose)), st)
                _cabi.call("lv_gemm_tf32_f32", _cabi.ptr(ybuf), K, _cabi.ptr(wt), ldb, _cabi.ptr(bias), int(bias_div), _cabi.ptr(out[lo:lo + n]),
                           Nout, n, Nout, K, st)
        ctx.save_for_backward(a_c, s_c, weight)
        ctx.meta = (N, lmax, C, K, Nout, int(bias_div), bool(transpose), chunk, bias is not None)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        a_c, s_c, weight = ctx.saved_tensors
        N, lmax, C, K, Nout, bias_div, transpose, chunk, has_bias = ctx.meta
        dev = a_c.device
        if gout is None:
            return (None,) * 8
        g = gout.contiguous()
        gang = torch.empty((N, 3), dtype=torch.float32, device=dev)
        gitem = torch.zeros_like(s_c)
        gw = torch.zeros((K, Nout), dtype=torch.float32, device=dev)
        ybuf = torch.empty((chunk, K), dtype=torch.float32, device=dev)
        gybuf = torch.empty((chunk, K), dtype=torch.float32, device=dev)
        w_r = round_tf32(weight) if Nout % 4 == 0 else None
        with _on(dev):
            st = _stream()
            nws = _cabi.lib().lv_wigner_bwd_workspace_floats(chunk, 0, lmax, C)
            ws = torch.empty(max(nws, 1), dtype=torch.float32, device=dev)
            for lo in range(0, N, chunk):
                n = min(chunk, N - lo)
                gc = g[lo:lo + n]
                _cabi.call("lv_wigner_apply_fwd_f32", _cabi.ptr(a_c[lo:lo + n]), _cabi.ptr(s_c), _cabi.ptr(ybuf), n, 0, lmax, C, 1, int(transpose), st)
                prev_tf32 = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = bool(ACTION_GEMM_WGRAD_TF32)
                try:
                    gw.addmm_(ybuf[:n].t(), gc)                              # wgrad (cuBLAS), y from L2
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = prev_tf32
                # dgrad on the tensor cores too: g_y = g_out (n, Nout) @ W^T -- W (K, Nout) is already the K-major "Bt" of this
                # product; g_y stays in L2 for the Wigner backward
                if w_r is not None:
                    _cabi.call("lv_gemm_tf32_f32", _cabi.ptr(gc), Nout, _cabi.ptr(w_r), Nout, None, 1, _cabi.ptr(gybuf), K, n, K, Nout, st)
                else:                                                       # weight rows not 16-byte aligned for TMA: cuBLAS
                    torch.matmul(gc, weight.t(), out=gybuf[:n])
                _cabi.call("lv_wigner_apply_bwd_f32", _cabi.ptr(a_c[lo:lo + n]), _cabi.ptr(s_c), _cabi.ptr(gybuf), _cabi.ptr(gang[lo:lo + n]),
                           _cabi.ptr(gitem), _cabi.ptr(ws), nws, n, 0, lmax, C, 3, int(transpose), st)
        gbias = g.view(N, Nout // bias_div, bias_div).sum((0, 2)) if has_bias else None
        return gang, gitem, gw, gbias, None, None, None, None
