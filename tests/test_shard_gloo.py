"""world_size-2 gloo run of the N>1 path on CPU: the object partition + one sum-reduce of the mix gives the
same track as the unsharded render.  (No GPU here, so each rank renders its shard with the oracle; the
partition and the collective are the code bench.py runs over NCCL.)"""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_obj, n_modes, n_buf, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from openpbso_b200 import synth
    from openpbso_b200.shard import shard_range, reduce_mix
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.batch_workload(n_obj, n_modes, n_buf, 77)
    lo, hi = shard_range(n_obj, world, rank)
    mix = np.zeros(n_buf * 256)
    if hi > lo:
        orc.batch_render(synth.H, w["a"][lo:hi], w["b"][lo:hi], w["space"][lo:hi], w["trans"][lo:hi], w["imp_buf"][lo:hi], 256, n_buf, mix)
    t = torch.from_numpy(mix)
    reduce_mix(t, 0)
    if rank == 0:
        np.save(os.path.join(out_dir, "mix.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_tile_the_objects():
    sys.path.insert(0, ROOT)
    from openpbso_b200.shard import shard_range
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= (n + world - 1) // world


@pytest.mark.parametrize("n_obj", [5, 8])
def test_two_rank_gloo_reduce_equals_unsharded(orc, tmp_path, n_obj):
    import torch.multiprocessing as mp
    from openpbso_b200 import synth
    n_modes, n_buf = 40, 12
    port = 29650 + n_obj
    mp.spawn(_worker, args=(2, port, n_obj, n_modes, n_buf, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "mix.npy"))
    w = synth.batch_workload(n_obj, n_modes, n_buf, 77)
    ref = orc.batch_render(synth.H, w["a"], w["b"], w["space"], w["trans"], w["imp_buf"], 256, n_buf)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
