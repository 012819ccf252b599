"""ctypes driver for the C oracle + the composition of the hot path on top of it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy in, numpy out; the weight
container is a dict keyed exactly like the reference's ``state_dict()``.
"""
import ctypes as C

import numpy as np

from .build import build

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        f32p, i64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        _lib.oracle_vq_forward_f32.argtypes = [f32p, f32p, C.c_int64, C.c_int, C.c_int, i64p,
                                               f32p, C.POINTER(C.c_double), i32p, f32p]
        _lib.oracle_vq_finish_f32.argtypes = [C.c_double, i32p, C.c_int64, C.c_int, C.c_int,
                                              C.c_float, f32p, f32p]
        conv_args = [f32p, f32p, f32p, f32p] + [C.c_int] * 9
        _lib.oracle_conv2d_f32.argtypes = conv_args
        _lib.oracle_conv_transpose2d_f32.argtypes = conv_args
        for fn in (_lib.oracle_vq_forward_f32, _lib.oracle_vq_finish_f32,
                   _lib.oracle_conv2d_f32, _lib.oracle_conv_transpose2d_f32):
            fn.restype = None
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty))


def vq_rows(z_rows, codebook, beta=0.25):
    """quantizer.py:46-71 on already-flattened rows.  Returns a dict."""
    z = _f32(z_rows)
    E = _f32(codebook)
    N, D = z.shape
    K = E.shape[0]
    assert E.shape[1] == D
    idx = np.empty(N, np.int64)
    zq = np.empty_like(z)
    hist = np.zeros(K, np.int32)
    dmin = np.empty(N, np.float32)
    sse = C.c_double(0.0)
    _L().oracle_vq_forward_f32(_p(z), _p(E), N, K, D, _p(idx, C.c_int64), _p(zq),
                               C.byref(sse), _p(hist, C.c_int32), _p(dmin))
    loss = C.c_float(0)
    perp = C.c_float(0)
    _L().oracle_vq_finish_f32(sse.value, _p(hist, C.c_int32), N, K, D, beta,
                              C.byref(loss), C.byref(perp))
    return dict(idx=idx, zq=zq, sse=sse.value, hist=hist, dmin=dmin,
                loss=np.float32(loss.value), perplexity=np.float32(perp.value))


def vq_nchw(z_nchw, codebook, beta=0.25):
    """VectorQuantizer.forward (quantizer.py:29-76) on a (B,D,H,W) tensor."""
    z = _f32(z_nchw)
    B, D, H, W = z.shape
    rows = np.ascontiguousarray(z.transpose(0, 2, 3, 1)).reshape(-1, D)   # :45-46
    r = vq_rows(rows, codebook, beta)
    r["zq_nchw"] = np.ascontiguousarray(r["zq"].reshape(B, H, W, D).transpose(0, 3, 1, 2))  # :74
    r["idx"] = r["idx"].reshape(-1, 1)                                     # :54 unsqueeze(1)
    return r


def conv2d(x, w, b, stride, pad):
    x, w = _f32(x), _f32(w)
    B, Cin, H, W = x.shape
    Cout, Cin2, kh, kw = w.shape
    assert Cin == Cin2
    OH = (H + 2 * pad - kh) // stride + 1
    OW = (W + 2 * pad - kw) // stride + 1
    y = np.empty((B, Cout, OH, OW), np.float32)
    bp = _p(_f32(b)) if b is not None else None
    _L().oracle_conv2d_f32(_p(x), _p(w), bp, _p(y), B, Cin, H, W, Cout, kh, kw, stride, pad)
    return y


def conv_transpose2d(x, w, b, stride, pad):
    x, w = _f32(x), _f32(w)
    B, Cin, H, W = x.shape
    Cin2, Cout, kh, kw = w.shape
    assert Cin == Cin2
    OH = (H - 1) * stride - 2 * pad + kh
    OW = (W - 1) * stride - 2 * pad + kw
    y = np.empty((B, Cout, OH, OW), np.float32)
    bp = _p(_f32(b)) if b is not None else None
    _L().oracle_conv_transpose2d_f32(_p(x), _p(w), bp, _p(y), B, Cin, H, W, Cout, kh, kw,
                                     stride, pad)
    return y


def relu(x):
    return np.maximum(x, np.float32(0))


def residual_stack(x, w1, w2, n_layers):
    """residual.py:41-51 with the quirks of SURVEY 3.3: ONE (w1,w2) pair applied
    n times (Q1); the in-place ReLU makes each layer relu(x) + f(relu(x)) (Q2);
    final F.relu (Q3)."""
    for _ in range(n_layers):
        r = relu(x)                                   # residual.py:19 (in place on x)
        h = relu(conv2d(r, w1, None, 1, 1))           # :20-22
        x = r + conv2d(h, w2, None, 1, 0)             # :23-24, :28
    return relu(x)                                    # :50


def encoder(x, sd, n_res_layers, prefix="encoder.conv_stack."):
    """encoder.py:28-43."""
    h = relu(conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], 2, 1))
    h = relu(conv2d(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"], 2, 1))
    h = conv2d(h, sd[prefix + "4.weight"], sd[prefix + "4.bias"], 1, 1)
    return residual_stack(h, sd[prefix + "5.stack.0.res_block.1.weight"],
                          sd[prefix + "5.stack.0.res_block.3.weight"], n_res_layers) \
        if n_res_layers > 0 else relu(h)


def decoder(z, sd, n_res_layers, prefix="decoder.inverse_conv_stack."):
    """decoder.py:27-39."""
    h = conv_transpose2d(z, sd[prefix + "0.weight"], sd[prefix + "0.bias"], 1, 1)
    if n_res_layers > 0:
        h = residual_stack(h, sd[prefix + "1.stack.0.res_block.1.weight"],
                           sd[prefix + "1.stack.0.res_block.3.weight"], n_res_layers)
    else:
        h = relu(h)
    h = relu(conv_transpose2d(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"], 2, 1))
    return conv_transpose2d(h, sd[prefix + "4.weight"], sd[prefix + "4.bias"], 2, 1)


def vqvae_forward(x, sd, n_res_layers, beta=0.25):
    """VQVAE.forward (vqvae.py:29-44); returns every intermediate for parity tests."""
    sd = {k: np.asarray(v, np.float32) for k, v in sd.items()}
    enc = encoder(x, sd, n_res_layers)                                         # :31
    z_e = conv2d(enc, sd["pre_quantization_conv.weight"],
                 sd["pre_quantization_conv.bias"], 1, 0)                       # :33
    vq = vq_nchw(z_e, sd["vector_quantization.embedding.weight"], beta)        # :34
    x_hat = decoder(vq["zq_nchw"], sd, n_res_layers)                           # :36
    return dict(enc=enc, z_e=z_e, idx=vq["idx"], z_q=vq["zq_nchw"], x_hat=x_hat,
                loss=vq["loss"], perplexity=vq["perplexity"], hist=vq["hist"], sse=vq["sse"])
