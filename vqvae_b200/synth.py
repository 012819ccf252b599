"""Deterministic synthetic weights and images keyed like the reference's state_dict.

No arithmetic of the path lives here -- only seeded random tensors -- so the product benchmark, the tests and the
golden generator all share it (oracle/weights.py re-exports it).  numpy's MT19937 stream is stable across platforms and
versions, so the same (hyper-params, seed) gives the same weights in the authoring
container (where the golden vectors are made from the unmodified reference) and on
the GPU box (where the CUDA path is checked against them).  Scales follow PyTorch's
default Conv init (U(+-1/sqrt(fan_in))) and quantizer.py:27 (U(+-1/K)).
"""
import re

import numpy as np


def state_dict_shapes(h_dim, res_h_dim, n_res_layers, n_embeddings, embedding_dim):
    """(key, shape, fan_in) in the reference's state_dict order.  The ResidualStack
    aliases ONE layer n times (residual.py:44-45), so stack.i.* keys repeat it."""
    h, r, K, D = h_dim, res_h_dim, n_embeddings, embedding_dim
    out = []
    e = "encoder.conv_stack."
    out += [(e + "0.weight", (h // 2, 3, 4, 4), 3 * 16), (e + "0.bias", (h // 2,), 3 * 16),
            (e + "2.weight", (h, h // 2, 4, 4), (h // 2) * 16), (e + "2.bias", (h,), (h // 2) * 16),
            (e + "4.weight", (h, h, 3, 3), h * 9), (e + "4.bias", (h,), h * 9)]
    for i in range(n_res_layers):
        out += [(e + f"5.stack.{i}.res_block.1.weight", (r, h, 3, 3), h * 9),
                (e + f"5.stack.{i}.res_block.3.weight", (h, r, 1, 1), r)]
    out += [("pre_quantization_conv.weight", (D, h, 1, 1), h),
            ("pre_quantization_conv.bias", (D,), h),
            ("vector_quantization.embedding.weight", (K, D), None)]
    d = "decoder.inverse_conv_stack."
    # ConvTranspose2d weights are (Cin, Cout, kh, kw); torch's fan_in uses dim 1.
    out += [(d + "0.weight", (D, h, 3, 3), h * 9), (d + "0.bias", (h,), h * 9)]
    for i in range(n_res_layers):
        out += [(d + f"1.stack.{i}.res_block.1.weight", (r, h, 3, 3), h * 9),
                (d + f"1.stack.{i}.res_block.3.weight", (h, r, 1, 1), r)]
    out += [(d + "2.weight", (h, h // 2, 4, 4), (h // 2) * 16), (d + "2.bias", (h // 2,), (h // 2) * 16),
            (d + "4.weight", (h // 2, 3, 4, 4), 3 * 16), (d + "4.bias", (3,), 3 * 16)]
    return out


def make_state_dict(h_dim=128, res_h_dim=32, n_res_layers=2, n_embeddings=512,
                    embedding_dim=64, seed=0, codebook="default", codebook_scale=1.0):
    """numpy state dict.  codebook: "default" = U(+-1/K) (near-tie stress);
    "normal" = N(0, codebook_scale^2) (trained-like, wide code usage)."""
    rng = np.random.RandomState(seed)
    sd = {}
    for key, shape, fan_in in state_dict_shapes(h_dim, res_h_dim, n_res_layers,
                                                n_embeddings, embedding_dim):
        alias = re.sub(r"\.stack\.\d+\.", ".stack.0.", key)
        if alias != key:
            sd[key] = sd[alias]          # same array object, like the shared layer
            continue
        if fan_in is None:
            if codebook == "default":
                b = 1.0 / n_embeddings
                sd[key] = rng.uniform(-b, b, size=shape).astype(np.float32)
            else:
                sd[key] = (rng.standard_normal(size=shape) * codebook_scale).astype(np.float32)
        else:
            b = 1.0 / np.sqrt(fan_in)
            sd[key] = rng.uniform(-b, b, size=shape).astype(np.float32)
    return sd


def make_images(batch, size, seed=1):
    """x = 2*rand - 1, the range CIFAR gets after Normalize(0.5, 0.5) (utils.py:13-17)."""
    rng = np.random.RandomState(seed)
    return (2.0 * rng.random_sample((batch, 3, size, size)) - 1.0).astype(np.float32)
