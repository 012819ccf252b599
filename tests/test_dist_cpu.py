"""N>1 host logic on CPU: world_size-2 gloo run of the batch-shard plumbing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cref
from tests.helpers import make_vq_inputs, VQ_CASES


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqvae_b200.dist import reduce_vq_stats, shard_bounds
    z, E = make_vq_inputs(**VQ_CASES["vq_k512_d64"])
    B, D = z.shape[0], z.shape[1]
    lo, hi = shard_bounds(B, rank, world)
    r = cref.vq_nchw(z[lo:hi], E)                      # this rank's shard through the oracle
    hist, sse = reduce_vq_stats(torch.from_numpy(r["hist"]), torch.tensor([r["sse"]], dtype=torch.float64))
    idx_all = [torch.zeros_like(torch.from_numpy(r["idx"])) for _ in range(world)]
    dist.all_gather(idx_all, torch.from_numpy(r["idx"]))
    if rank == 0:
        np.savez(out, hist=hist.numpy(), sse=sse.numpy(), idx=torch.cat(idx_all).numpy())
    dist.destroy_process_group()


def test_sharded_stats_equal_single_process(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    z, E = make_vq_inputs(**VQ_CASES["vq_k512_d64"])
    full = cref.vq_nchw(z, E)
    assert np.array_equal(got["hist"], full["hist"])                 # exact counts
    np.testing.assert_allclose(got["sse"][0], full["sse"], rtol=1e-12)
    assert np.array_equal(got["idx"], full["idx"])                   # rank-major = reference row order


def test_shard_bounds():
    from vqvae_b200.dist import shard_bounds
    assert [shard_bounds(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    try:
        shard_bounds(10, 0, 4)
    except ValueError:
        pass
    else:
        raise AssertionError("uneven shards must be rejected")
