"""The product's own multi-GPU path on real devices (skips below 2 GPUs): one process per GPU, pbso_comm_* (NCCL
from the C ABI) shards the units and reduces the audio; the sharded mix must equal the unsharded render.
Units are whole sound objects or mode blocks of ONE large object (the reference's hot loop is a
sum over modes, modal_solver.h:261-272, so a mode block is an independent unit with the same impulse times)."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, mode, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import openpbso_b200 as pbso
    from openpbso_b200 import synth
    pbso.set_device(rank); torch.cuda.set_device(rank)
    comm = pbso.Comm(world, rank, uid)
    n_buf = 150
    if mode == "objects":
        n_obj, n_modes = 37, 96
        w = synth.batch_workload(n_obj, n_modes, n_buf, 91, "high_damping", first_second_bufs=100)
        a, b, trans, space, imp = w["a"], w["b"], w["trans"], w["space"], w["imp_buf"]
    else:
        # one object with 4096 modes cut into 16 blocks of 256 modes: every block is hit by the same impulse
        n_modes_total, blk = 4096, 256
        w = synth.batch_workload(1, n_modes_total, n_buf, 92, "low_damping", first_second_bufs=60)
        n_obj, n_modes = n_modes_total // blk, blk
        a, b, trans, space = (w[k].reshape(n_obj, n_modes) for k in ("a", "b", "trans", "space"))
        imp = np.repeat(w["imp_buf"], n_obj)
    lo, hi = comm.shard(n_obj)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    mix = torch.zeros(n_buf * 256, dtype=torch.float64, device="cuda")
    for prec in (pbso.PREC_F64, pbso.PREC_TC3X):
        if hi > lo:
            br = pbso.BatchRenderer(synth.H, a[lo:hi], b[lo:hi]); br.set_stream(stream.cuda_stream)
            br.set_transfer(trans[lo:hi]); br.set_impulses(np.arange(hi - lo), imp[lo:hi], space[lo:hi])
            br.render_mix_device(256, n_buf, mix.data_ptr(), prec)
        else:
            mix.zero_()
        comm.reduce_audio(mix.data_ptr(), mix.numel(), 0, stream.cuda_stream)
        torch.cuda.synchronize()
        if rank == 0:
            full = pbso.BatchRenderer(synth.H, a, b); full.set_transfer(trans); full.set_impulses(np.arange(n_obj), imp, space)
            ref = full.render_mix(256, n_buf, pbso.PREC_F64)
            got = mix.cpu().numpy()
            tol = 1e-12 if prec == pbso.PREC_F64 else 1e-6
            err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
            np.save(os.path.join(out_dir, "err_%s_%d.npy" % (mode, prec)), np.array([err, tol]))
    comm.close()


@pytest.mark.parametrize("mode", ["objects", "mode_blocks"])
def test_sharded_render_and_nccl_reduce_equal_the_unsharded_mix(pbso, tmp_path, mode):
    if pbso.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    uid = pbso.Comm.unique_id()
    mp.spawn(_worker, args=(2, uid, mode, str(tmp_path)), nprocs=2, join=True)
    for prec in (pbso.PREC_F64, pbso.PREC_TC3X):
        err, tol = np.load(os.path.join(str(tmp_path), "err_%s_%d.npy" % (mode, prec)))
        assert err <= tol, (mode, prec, err)


def test_single_rank_comm_is_a_no_op(pbso):
    import torch
    c = pbso.Comm(1, 0)
    assert c.shard(10) == (0, 10)
    x = torch.arange(8, dtype=torch.float64, device="cuda")
    c.reduce_audio(x.data_ptr(), 8, 0)
    torch.cuda.synchronize()
    assert x.cpu().tolist() == list(range(8))
    c.close()


def test_render_tool_batch_mode_over_two_gpus(tmp_path):
    """tools/pbso_render -batch -gpus 2: one rank (thread + device) per mode block, pbso_comm_reduce_audio_host lands the
    track on rank 0 -- the same script rendered on one device must give the same waveform."""
    import subprocess
    sys.path.insert(0, ROOT)
    import openpbso_b200 as pbso
    if pbso.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_headless import _write_object, SCRIPT_LONG, GXX, LIBDIR
    from conftest import assert_waveform_parity
    exe = str(tmp_path / "pbso_render")
    r = subprocess.run(GXX + [os.path.join(ROOT, "tools", "pbso_render.cpp"), "-L" + LIBDIR, "-lpbso_b200", "-Wl,-rpath," + LIBDIR, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = str(tmp_path); _write_object(d, "bell", orc)
    script = os.path.join(d, "l.txt"); open(script, "w").write(SCRIPT_LONG)
    outs = []
    for gpus in (1, 2):
        raw = os.path.join(d, "g%d.f64" % gpus)
        r = subprocess.run([exe, "-d", d, "-script", script, "-buf", "513", "-raw", raw, "-batch", "-gpus", str(gpus), "-block", "16"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "on %d device(s)" % gpus in r.stdout
        outs.append(np.fromfile(raw))
    assert outs[0].size == outs[1].size and np.abs(outs[0]).max() > 0
    assert_waveform_parity(outs[1], outs[0])
