"""Batch-sharded inference plumbing: the path shards by images with no
data-path collective; the only cross-sample quantities of VQVAE.forward are the SSE of
the embedding loss and the code histogram of the perplexity (quantizer.py:63-64, :70-71).

Pure torch.distributed on raw tensors, so it runs on NCCL (GPU) and on gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous, equal-sized image shard [lo, hi) of rank `rank` (rank-major order = the
    reference's row order).  The global batch must divide evenly so that every rank can
    derive the global row count without a host sync."""
    if batch % world:
        raise ValueError(f"global batch {batch} is not divisible by world size {world}")
    per = batch // world
    return rank * per, (rank + 1) * per


def reduce_vq_stats(hist: torch.Tensor, sse: torch.Tensor, group=None):
    """All-reduce (sum) the int32 code histogram (K,) and the float64 SSE (1,) in ONE
    collective (counts are exact in float64 up to 2^53).  Returns (hist int32, sse f64)."""
    stats = torch.cat([hist.to(torch.float64), sse.to(torch.float64).reshape(1)])
    dist.all_reduce(stats, group=group)
    return stats[:-1].round().to(torch.int32), stats[-1:].contiguous()
