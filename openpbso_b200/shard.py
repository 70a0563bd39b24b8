"""Multi-GPU decomposition of the offline batch: sound objects are independent, so each
rank renders a contiguous block of objects and the only exchange is one sum-reduce of the mixed-down track.
Used by bench.py (NCCL) and by the world_size-2 gloo test (tests/test_shard_gloo.py)."""


def shard_range(n_obj, world, rank):
    """Contiguous block [lo, hi) of objects for `rank`; the blocks tile [0, n_obj) without overlap."""
    per = (n_obj + world - 1) // world
    lo = min(rank * per, n_obj)
    return lo, min(lo + per, n_obj)


def reduce_mix(mix, dst=0):
    """Sum the per-rank partial mixes onto `dst` (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM)
    return mix
