"""The hot path restated with the torch CPU ops the reference itself dispatches to.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the ``cpu_baseline`` /
``--impl reference`` arm of bench.py ("kind": "port"): /root/reference does not
exist on the GPU box, so the reference's forward is restated here functionally on a
state dict.  It runs on the same oneDNN / MKL back ends as the reference, hence the
same speed class, and is pinned against the unmodified reference by
tests/test_oracle.py (golden vectors made by oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F


def residual_stack(x, w1, w2, n_layers):
    # residual.py:44-51 -- one shared layer applied n times; the
    # in-place ReLU means the skip branch carries relu(x) (Q2).
    for _ in range(n_layers):
        r = F.relu(x)
        x = r + F.conv2d(F.relu(F.conv2d(r, w1, None, 1, 1)), w2, None, 1, 0)
    return F.relu(x)


def encoder(x, sd, n_res, p="encoder.conv_stack."):
    # encoder.py:28-40
    h = F.relu(F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], 2, 1))
    h = F.relu(F.conv2d(h, sd[p + "2.weight"], sd[p + "2.bias"], 2, 1))
    h = F.conv2d(h, sd[p + "4.weight"], sd[p + "4.bias"], 1, 1)
    if n_res == 0:
        return F.relu(h)
    return residual_stack(h, sd[p + "5.stack.0.res_block.1.weight"],
                          sd[p + "5.stack.0.res_block.3.weight"], n_res)


def decoder(z, sd, n_res, p="decoder.inverse_conv_stack."):
    # decoder.py:27-36
    h = F.conv_transpose2d(z, sd[p + "0.weight"], sd[p + "0.bias"], 1, 1)
    if n_res == 0:
        h = F.relu(h)
    else:
        h = residual_stack(h, sd[p + "1.stack.0.res_block.1.weight"],
                           sd[p + "1.stack.0.res_block.3.weight"], n_res)
    h = F.relu(F.conv_transpose2d(h, sd[p + "2.weight"], sd[p + "2.bias"], 2, 1))
    return F.conv_transpose2d(h, sd[p + "4.weight"], sd[p + "4.bias"], 2, 1)


def vector_quantizer(z, E, beta):
    # quantizer.py:45-76, op for op (including the dense one-hot and the second GEMM,
    # which is what the reference spends its time on).
    z = z.permute(0, 2, 3, 1).contiguous()
    zf = z.view(-1, E.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1) \
        - 2 * torch.matmul(zf, E.t())
    idx = torch.argmin(d, dim=1).unsqueeze(1)
    onehot = torch.zeros(idx.shape[0], E.shape[0])
    onehot.scatter_(1, idx, 1)
    z_q = torch.matmul(onehot, E).view(z.shape)
    loss = torch.mean((z_q - z) ** 2) + beta * torch.mean((z_q - z) ** 2)
    z_q = z + (z_q - z)
    e_mean = torch.mean(onehot, dim=0)
    perplexity = torch.exp(-torch.sum(e_mean * torch.log(e_mean + 1e-10)))
    return loss, z_q.permute(0, 3, 1, 2).contiguous(), perplexity, onehot, idx


@torch.no_grad()
def vqvae_forward(x, sd, n_res, beta=0.25, intermediates=False):
    # vqvae.py:29-44
    z_e = encoder(x, sd, n_res)
    z_e = F.conv2d(z_e, sd["pre_quantization_conv.weight"], sd["pre_quantization_conv.bias"])
    loss, z_q, perplexity, _, idx = vector_quantizer(
        z_e, sd["vector_quantization.embedding.weight"], beta)
    x_hat = decoder(z_q, sd, n_res)
    if intermediates:
        return dict(z_e=z_e, idx=idx, z_q=z_q, x_hat=x_hat, loss=loss, perplexity=perplexity)
    return loss, x_hat, perplexity
