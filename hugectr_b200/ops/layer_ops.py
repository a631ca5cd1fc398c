"""ctypes wrappers for csrc/layers.cu (native forward / backward of the long-tail dense layers).

CUDA only: the CPU path of these layers is their PyTorch formulation (``TorchLayer``), which is also the
numerics oracle of the GPU tests."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native
from . import dense as D

_lib = None


class CBmmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("Z1", C.c_int),
                ("a_z0", C.c_longlong), ("a_z1", C.c_longlong), ("a_m", C.c_longlong), ("a_k", C.c_longlong),
                ("b_z0", C.c_longlong), ("b_z1", C.c_longlong), ("b_k", C.c_longlong), ("b_n", C.c_longlong),
                ("c_z0", C.c_longlong), ("c_z1", C.c_longlong), ("c_m", C.c_longlong), ("c_n", C.c_longlong),
                ("alpha", C.c_float), ("accumulate", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        l = _native.cuda_lib()
        vp, ll, i, f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
        l.hctr_softmax_fwd.argtypes = [vp, vp, vp, ll, i, i, vp]
        l.hctr_softmax_bwd.argtypes = [vp, vp, vp, vp, ll, i, i, vp]
        l.hctr_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, vp, ll, i, f, i, vp]
        l.hctr_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ll, i, i, vp]
        l.hctr_colreduce2.argtypes = [vp, vp, vp, vp, vp, vp, ll, i, f, i, i, vp]
        l.hctr_bn_finalize.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, ll, f, f, vp]
        l.hctr_colwise.argtypes = [vp, vp, vp, vp, vp, vp, vp, ll, i, f, i, i, vp]
        l.hctr_bn_bwd_dx.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ll, i, i, vp]
        l.hctr_fm_order2.argtypes = [vp, vp, vp, ll, i, i, i, i, vp]
        l.hctr_weight_mul_fwd.argtypes = [vp, vp, vp, ll, i, i, i, vp]
        l.hctr_weight_mul_bwd.argtypes = [vp, vp, vp, vp, vp, ll, i, i, i, vp]
        l.hctr_reduce_mid.argtypes = [vp, vp, ll, i, ll, f, i, i, vp]
        l.hctr_copy4d.argtypes = [vp, vp, C.POINTER(ll), C.POINTER(ll), C.POINTER(ll), i, i, vp]
        l.hctr_bmm.argtypes = [vp, vp, vp, C.POINTER(CBmmDesc), i, i, i, i, vp]
        l.hctr_gru_gate.argtypes = [vp, vp, vp, vp, vp, vp, vp, ll, i, i, vp]
        for n in ("hctr_softmax_fwd", "hctr_softmax_bwd", "hctr_layernorm_fwd", "hctr_layernorm_bwd",
                  "hctr_colreduce2", "hctr_bn_finalize", "hctr_colwise", "hctr_bn_bwd_dx", "hctr_fm_order2",
                  "hctr_weight_mul_fwd", "hctr_weight_mul_bwd", "hctr_reduce_mid", "hctr_copy4d", "hctr_bmm",
                  "hctr_gru_gate"):
            getattr(l, n).restype = i
        if l.hctr_abi_size_bmm() != C.sizeof(CBmmDesc):
            raise RuntimeError("libhctr_cuda.so BmmDesc layout differs from the python mirror: rebuild")
        _lib = l
    return _lib


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _bf(t):
    return int(t.dtype == torch.bfloat16)


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(rc, name, n=1):
    if rc:
        raise RuntimeError(f"{name} failed rc={rc}")
    D._count(n)


def ok(*ts) -> bool:
    """native kernels apply: CUDA, fp32 / bf16, contiguous, one activation dtype"""
    ts = [t for t in ts if t is not None]
    return bool(ts) and all(t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16)
                            for t in ts) and len({t.dtype for t in ts}) == 1


def softmax_fwd(x, mask, y):
    n = x.shape[-1]
    _chk(lib().hctr_softmax_fwd(x.data_ptr(), _p(mask), y.data_ptr(), x.numel() // n, n, _bf(x), _st(x)), "softmax_fwd")


def softmax_bwd(dy, y, mask, dx):
    n = y.shape[-1]
    _chk(lib().hctr_softmax_bwd(dy.data_ptr(), y.data_ptr(), _p(mask), dx.data_ptr(), y.numel() // n, n, _bf(y),
                                _st(y)), "softmax_bwd")


def layernorm_fwd(x, gamma, beta, y, mean, rstd, eps):
    n = x.shape[-1]
    _chk(lib().hctr_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                  rstd.data_ptr(), x.numel() // n, n, float(eps), _bf(x), _st(x)), "layernorm_fwd")


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta):
    """dx (may be None) overwritten; dgamma / dbeta (fp32 [n]) accumulated"""
    n = x.shape[-1]
    _chk(lib().hctr_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                  _p(dx), dgamma.data_ptr(), dbeta.data_ptr(), x.numel() // n, n, _bf(x), _st(x)),
         "layernorm_bwd", 2)


def colreduce2(x, dy, m, s, out0, out1, mode, alpha=0.0):
    cols = x.shape[-1]
    _chk(lib().hctr_colreduce2(x.data_ptr(), _p(dy), _p(m), _p(s), out0.data_ptr(), out1.data_ptr(),
                               x.numel() // cols, cols, float(alpha), mode, _bf(x), _st(x)), "colreduce2")


def bn_finalize(sum_, sumsq, mean, rstd, var_out, run_mean, run_var, rows, eps, momentum):
    _chk(lib().hctr_bn_finalize(sum_.data_ptr(), sumsq.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _p(var_out),
                                _p(run_mean), _p(run_var), sum_.numel(), rows, float(eps), float(momentum),
                                _st(sum_)), "bn_finalize")


def colwise(x, dy, m, s, p0, p1, out, op, alpha=0.0):
    cols = x.shape[-1]
    _chk(lib().hctr_colwise(x.data_ptr(), _p(dy), _p(m), _p(s), _p(p0), _p(p1), out.data_ptr(), x.numel() // cols,
                            cols, float(alpha), op, _bf(x), _st(x)), "colwise")


def bn_bwd_dx(x, dy, mean, rstd, gamma, dbeta, dgamma, dx):
    cols = x.shape[-1]
    _chk(lib().hctr_bn_bwd_dx(x.data_ptr(), dy.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                              dbeta.data_ptr(), dgamma.data_ptr(), dx.data_ptr(), x.numel() // cols, cols, _bf(x),
                              _st(x)), "bn_bwd_dx")


def fm_order2(x, dy, out, b, S, Dm, backward):
    _chk(lib().hctr_fm_order2(x.data_ptr(), _p(dy), out.data_ptr(), b, S, Dm, int(backward), _bf(x), _st(x)),
         "fm_order2")


def weight_mul_fwd(x, w, y, b, S, V):
    _chk(lib().hctr_weight_mul_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), b, S, V, _bf(x), _st(x)), "weight_mul_fwd")


def weight_mul_bwd(dy, x, w, dx, dw, b, S, V):
    _chk(lib().hctr_weight_mul_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), _p(dx), dw.data_ptr(), b, S, V,
                                   _bf(x), _st(x)), "weight_mul_bwd", 2)


def reduce_mid(inp, out, outer, R, inner, scale, backward):
    _chk(lib().hctr_reduce_mid(inp.data_ptr(), out.data_ptr(), outer, R, inner, float(scale), int(backward),
                               _bf(inp), _st(inp)), "reduce_mid")


def copy4d(src, dst, dims, sstr, dstr, accumulate=False, src_off=0, dst_off=0):
    """dst[dst_off + sum i*dstr] (+)= src[src_off + sum i*sstr] over a 4-D index space (element strides;
    a source stride of 0 broadcasts)"""
    ll4 = C.c_longlong * 4
    k = 4 - len(dims)
    d, ss, ds = list(dims) + [1] * k, list(sstr) + [0] * k, list(dstr) + [0] * k
    esz = src.element_size()
    _chk(lib().hctr_copy4d(src.data_ptr() + src_off * esz, dst.data_ptr() + dst_off * esz, ll4(*d), ll4(*ss),
                           ll4(*ds), int(accumulate), esz, _st(src)), "copy4d")


def bmm(a, b, c, M, N, K, Z0, Z1, a_str, b_str, c_str, alpha=1.0, accumulate=False):
    """C[z0, z1](m, n) (+)= alpha * sum_k A[z0, z1](m, k) B[z0, z1](k, n); *_str = (z0, z1, row, col) element
    strides of each operand (rows / cols as named in the formula)."""
    d = CBmmDesc(M, N, K, Z1, *a_str, *b_str, *c_str, float(alpha), int(accumulate))
    _chk(lib().hctr_bmm(a.data_ptr(), b.data_ptr(), c.data_ptr(), C.byref(d), Z0, _bf(a), _bf(b), _bf(c), _st(a)),
         "bmm")


def gru_gate_fwd(gi, gh, hprev, hnew, save, b, H):
    _chk(lib().hctr_gru_gate(gi.data_ptr(), gh.data_ptr(), hprev.data_ptr(), 0, hnew.data_ptr(), save.data_ptr(), 0,
                             b, H, 0, _st(gi)), "gru_gate_fwd")


def gru_gate_bwd(dh, save, gh, hprev, dgi, dgh, dhprev, b, H):
    _chk(lib().hctr_gru_gate(dh.data_ptr(), save.data_ptr(), gh.data_ptr(), hprev.data_ptr(), dgi.data_ptr(),
                             dgh.data_ptr(), dhprev.data_ptr(), b, H, 1, _st(dh)), "gru_gate_bwd")
