#!/usr/bin/env python
"""bench.py -- headline benchmark of the modal-synthesis path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): mode-samples/s (IIR + FFAT-weighted modal sum), whole job over N GPUs.
Workload: SURVEY.md 8(d) cfg5 -- offline batch, 4096 objects x 512 modes x 10 s of audio at 44.1 kHz
(1723 buffers of 256 samples), one PointForce per object in the first second, one static listener per
object (FFAT transfer vector resident), mixed down to one track.  Objects are block-partitioned over
the N ranks (strong scaling, total work fixed); the only exchange is one NCCL reduce (sum) of the
441 088-sample FP64 mix.  One "step" = one full render of that job.

The default path (--precision tc3x) is the tensor-core formulation of the synthesis (batch_tc.cu: pole-power
contraction, tcgen05 3xTF32); --precision f32_tiled selects the FP32-FMA pole-power kernel, f64 the reference
arithmetic.  The JSON line also carries: the roofline of the synthesis kernel (tensor pipe for tc3x, against
half the measured bf16 rate; FP32-FMA for f32_tiled, peak measured in the same run with an FMA
micro-benchmark; both report the algorithmic 8 FLOP per mode-sample figure against the FP32-FMA peak as the
north star asks), the CPU oracle timed on this box's host cores on a bounded sample
(`cpu_baseline`), the end-to-end number through the host-pointer C ABI (`e2e`), and the real-time
per-buffer latency of the second half of the metric (cfg2: 1024 modes, 256-sample buffers) under
`realtime`.

`--impl reference` times the reference's CPU path (the oracle port of modal_solver.h:181-276; the
reference itself cannot be built here) with all host threads on bounded samples.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBJ, N_MODES, BUF, N_BUF = 4096, 512, 256, 1723      # cfg5: 441 088 samples = 10.0 s
TC3X_DRAM_BYTES_PER_MODE_SAMPLE = 1.21867e9 / 9.2504e11   # dram__bytes_read + _write of k_batch_tc<2> on a full cfg5 launch (profiles/r2_k_batch_tc_final_metrics.txt)
FLOP_PER_MODE_SAMPLE = 8.0                              # 4 FMA: 3 in Step (modal_integrator.h:109-110) + 1 in the dot (modal_solver.h:267-269)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=N_OBJ, help="total objects (default cfg5: 4096)")
    ap.add_argument("--modes", type=int, default=N_MODES)
    ap.add_argument("--buffers", type=int, default=N_BUF)
    ap.add_argument("--precision", default="tc3x", choices=["tc3x", "f32_tiled", "f64"])
    ap.add_argument("--no-realtime", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the K3/K4/K5/K6 micro-benchmarks folded in under 'kernels'")
    ap.add_argument("--no-parity", action="store_true", help="skip the FP64 re-render of the job that 'parity' compares with")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_render_sample(n_threads, obj_per_thread, n_modes, n_buf, seed):
    """Each host thread renders `obj_per_thread` cfg5 objects for the full 10 s with the oracle's
    ModalSolver::step loop.  Returns (mode_samples, seconds)."""
    from oracle import oracle as orc
    from openpbso_b200 import synth
    orc.lib()
    n = n_threads * obj_per_thread
    w = synth.batch_workload(n, n_modes, n_buf, seed)
    mixes = [np.zeros(n_buf * BUF) for _ in range(n_threads)]

    def work(t):
        lo, hi = t * obj_per_thread, (t + 1) * obj_per_thread
        orc.batch_render(synth.H, w["a"][lo:hi], w["b"][lo:hi], w["space"][lo:hi], w["trans"][lo:hi],
                         w["imp_buf"][lo:hi], BUF, n_buf, mixes[t])
    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for th in threads: th.start()
    for th in threads: th.join()
    dt = time.perf_counter() - t0
    return float(n) * n_modes * n_buf * BUF, dt


def cpu_render_sample_ref(n_threads, n_modes, n_buf, seed):
    """The same loop run by the reference's OWN headers (oracle/_ref: ModalSolver<double,256>::step compiled in place
    against the Eigen shim), one object per host thread.  Returns (mode_samples, seconds) or None when _ref is absent.
    Reported beside the port on the same sample (same objects, same length); the faster of the two loops on a given host is
    the conservative CPU number -- in this container the shim's eager temporaries make the headers build 4-5x slower than
    the port, on the round-1 GPU box the two were equal."""
    from oracle import oracle as orc
    from openpbso_b200 import synth
    if orc.ref() is None:
        return None
    w = synth.batch_workload(n_threads, n_modes, n_buf, seed)

    def work(t):
        orc.ref_batch_render(synth.H, w["a"][t:t + 1], w["b"][t:t + 1], w["space"][t:t + 1], w["trans"][t:t + 1],
                             w["imp_buf"][t:t + 1], n_buf)
    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for th in threads: th.start()
    for th in threads: th.join()
    return float(n_threads) * n_modes * n_buf * BUF, time.perf_counter() - t0


def reference_headers_entry(cores, n_modes, n_buf):
    """The same sample as the port's (one object per host thread, full length), run by the reference's own headers."""
    r = cpu_render_sample_ref(cores, n_modes, n_buf, 1005)
    if r is None:
        return None
    return {"value": r[0] / r[1], "unit": "mode-samples/s", "cores": cores, "kind": "reference",
            "sample": "%d objects x %d modes x %d samples (1 object per host thread), the reference's own ModalSolver::step "
                      "compiled in place against the Eigen shim (oracle/_ref), %.1f s" % (cores, n_modes, n_buf * BUF, r[1])}


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return model, os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, cores = cpu_info()
    obj_per_thread = 1
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_render_sample(cores, 1, args.modes, max(args.buffers // 8, 1), 1)
    tot_ms = 0.0; tot_dt = 0.0
    for k in range(args.steps):
        ms, dt = cpu_render_sample(cores, obj_per_thread, args.modes, args.buffers, 1005 + k)
        tot_ms += ms; tot_dt += dt
    value = tot_ms / tot_dt
    sample = "%d objects x %d modes x %d samples per step (1 object per host thread), oracle port of ModalSolver::step" % (
        cores * obj_per_thread, args.modes, args.buffers * BUF)
    line = {
        "impl": "reference", "metric": "mode-samples/s (IIR+FFAT)", "value": value, "unit": "mode-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg5 offline batch (bounded sample): %s" % sample, "cpu": model},
        "cpu_baseline": {"value": value, "unit": "mode-samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "mode-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    rh = reference_headers_entry(cores, args.modes, args.buffers)
    if rh:
        line["cpu_baseline"]["reference_headers"] = rh
    emit(line)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML in a thread every 10 ms (the
    timed region of a multi-GPU run is only tens of milliseconds), `nvidia-smi -lms 100` if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index; self.rows = []; self.p = None; self.t_begin = None; self.t_end = None
        self.nvml = None; self.stop_flag = False; self.source = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml; self.source = "nvml, 10 ms period"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True); self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.p = None

    def _poll_nvml(self):
        n = self.nvml
        R = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
             ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                flags = ["Active" if mask & bit else "Not Active" for _, bit in R]
                self.rows.append((time.time(), [str(self.idx), sm, mx, pw, hex(mask)] + flags))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True; self.t.join(timeout=1)
        elif self.p:
            self.p.terminate()
            try: self.p.wait(timeout=2)
            except Exception: self.p.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        sm = []; mx = None; reasons = set(); power = []
        # samples taken inside the timed region; if the region was shorter than the sampling period, the
        # samples taken under the same load during warm-up are used and the fact is recorded
        inside = [r for (t, r) in self.rows if self.t_begin is not None and self.t_begin <= t <= (self.t_end or t)]
        window = "timed region" if inside else "warm-up + timed region (timed region shorter than the sampling period)"
        for r in (inside or [r for (_, r) in self.rows]):
            try:
                sm.append(float(r[1])); mx = float(r[2]) if r[2] is not None else mx; power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError, TypeError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "window": window, "source": self.source}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def realtime_latency(pbso, synth, n_buffers=10000):
    """cfg2: one 1024-mode object, 256-sample buffers, Bernoulli(0.12) impulse stream, host pointers in
    and out through pbso_render_buffer (H2D + kernel + D2H + sync per buffer); the listener jumps every 50
    buffers (computeTransfer on the device, K3 -> resident transfer table).  The jump is inside the timed call."""
    N = 1024
    mat = synth.MATERIALS["low_damping"]
    f = synth.mode_frequencies(N, 1002)
    a, b = synth.ab_from_material(f, mat)
    it = pbso.ModalIntegrator(N, synth.H, a, b)
    fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(f, 2000))
    rng = np.random.default_rng(1002)
    listeners = synth.listeners(n_buffers // 50 + 1, 1002)
    it.set_transfer_ffat(fm, listeners[:1])
    spaces = rng.standard_normal((64, N)); zero = np.zeros(N)
    tm_imp = np.zeros(BUF); tm_imp[0] = 1.0; tm_zero = np.zeros(BUF)
    hits = rng.random(n_buffers) < 0.12
    for _ in range(200):
        it.render_buffer(zero, tm_zero)
    # The timed loop calls the C ABI itself on preallocated buffers: the Python object wrapper allocates two arrays per
    # call, which shows up as a ~60 us hiccup on 1 % of the buffers (scripts/rt_latency_probe.py) that is not the
    # library's.  Arguments are marshalled before the clock starts only where a C caller would have them ready too.
    from openpbso_b200 import _capi as capi
    L_ = pbso.lib()
    y = np.empty(BUF); qn = np.empty(N)
    a_imp = [(it._h, capi.dp(spaces[j]), capi.dp(tm_imp), BUF, capi.dp(y), capi.dp(qn)) for j in range(64)]
    a_zero = (it._h, capi.dp(zero), capi.dp(tm_zero), BUF, capi.dp(y), capi.dp(qn))
    a_jump = [(it._h, fm._h, N, capi.dp(listeners[j:j + 1]), 1) for j in range(len(listeners))]
    lat = np.empty(n_buffers)
    for i in range(n_buffers):
        args = a_imp[i & 63] if hits[i] else a_zero
        t0 = time.perf_counter()
        if i % 50 == 0:
            rc = L_.pbso_integrator_set_transfer_ffat(*a_jump[i // 50])
        rc = L_.pbso_render_buffer(*args)
        lat[i] = time.perf_counter() - t0
        if rc:
            capi.check(rc)
    it.close()
    us = lat * 1e6
    return {"workload": "cfg2: 1024 modes, 1 listener (jumps every 50 buffers, FFAT re-evaluated on the device), 256-sample buffers, Bernoulli(0.12) PointForce stream, host in/out through the C ABI",
            "buffers": n_buffers, "p98_us": float(np.percentile(us, 98)), "p50_us": float(np.percentile(us, 50)), "p99_us": float(np.percentile(us, 99)),
            "p999_us": float(np.percentile(us, 99.9)), "max_us": float(us.max()),
            "mode_samples_per_s": float(N * BUF / np.mean(lat)), "dtype": "f64",
            "budget_us": 1e6 * BUF / synth.SAMPLE_RATE}


def moving_listeners_latency(pbso, synth, n_buffers=2000):
    """cfg4: 1024 modes, 64 listeners random-walking on the r = 5 sphere; every buffer: 64 positions H2D, K3 into the
    resident transfer table, one IIR pass rendering 64 outputs, 64 x 256 doubles D2H."""
    N, L = 1024, 64
    mat = synth.MATERIALS["low_damping"]
    f = synth.mode_frequencies(N, 1004)
    a, b = synth.ab_from_material(f, mat)
    it = pbso.ModalIntegrator(N, synth.H, a, b)
    fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(f, 2000))
    rng = np.random.default_rng(1004)
    pos = synth.listeners(L, 1004); pos *= 5.0 / np.linalg.norm(pos, axis=1, keepdims=True)
    spaces = rng.standard_normal((64, N)); zero = np.zeros(N)
    tm_imp = np.zeros(BUF); tm_imp[0] = 1.0; tm_zero = np.zeros(BUF)
    hits = rng.random(n_buffers) < 0.12
    steps = 0.02 * rng.standard_normal((64, L, 3))
    it.set_transfer_ffat(fm, pos)
    for _ in range(100):
        it.render_buffer(zero, tm_zero, want_qnorm=False)
    from openpbso_b200 import _capi as capi
    L_ = pbso.lib()
    y = np.empty((L, BUF))
    a_imp = [(it._h, capi.dp(spaces[j]), capi.dp(tm_imp), BUF, capi.dp(y), None) for j in range(64)]
    a_zero = (it._h, capi.dp(zero), capi.dp(tm_zero), BUF, capi.dp(y), None)
    lat = np.empty(n_buffers)
    for i in range(n_buffers):
        pos = pos + steps[i & 63]; pos *= 5.0 / np.linalg.norm(pos, axis=1, keepdims=True)
        pos = np.ascontiguousarray(pos)
        args = a_imp[i & 63] if hits[i] else a_zero
        a_pos = (it._h, fm._h, N, capi.dp(pos), L)
        t0 = time.perf_counter()
        rc = L_.pbso_integrator_set_transfer_ffat(*a_pos)
        rc = rc or L_.pbso_render_buffer(*args)
        lat[i] = time.perf_counter() - t0
        if rc:
            capi.check(rc)
    it.close()
    us = lat * 1e6
    return {"workload": "cfg4: 1024 modes, 64 moving listeners, FFAT re-evaluated every 256-sample buffer, host in/out through the C ABI",
            "buffers": n_buffers, "p50_us": float(np.percentile(us, 50)), "p99_us": float(np.percentile(us, 99)),
            "max_us": float(us.max()), "mode_samples_per_s": float(N * BUF / np.mean(lat)),
            "listener_mode_samples_per_s": float(N * BUF * L / np.mean(lat)), "dtype": "f64",
            "budget_us": 1e6 * BUF / synth.SAMPLE_RATE}


def contact_storm(pbso, synth, tf32_peak, n_buffers=300):
    """cfg3: 100 k impulses/s on 2048 modes x 20 000 vertices = 580 vertex impulses per 256-sample buffer, projected by the
    tensor-core GEMM (K5) and rendered (K1) without leaving the device: pbso_modes_storm_buffer with host pointers in and out.
    Parity: the same buffers through the FP64 sparse-gather path (the reference's arithmetic; tests/ pins it to the oracle)."""
    M, V, B, T = 2048, 20000, 580, BUF
    K = 3 * V
    mat = synth.MATERIALS["low_damping"]
    f = synth.mode_frequencies(M, 1003)
    a, b = synth.ab_from_material(f, mat)
    U = synth.mode_shapes(M, K, 1003)
    md = pbso.ModeShapes(U)
    rng = np.random.default_rng(1003)
    tr = np.abs(rng.standard_normal(M)) + 0.1
    vids = rng.integers(0, V, (16, B)).astype(np.int32)
    vns = np.stack([synth.unit_vectors(B, 3000 + i) for i in range(16)])
    res = {}
    ys = {}
    for name, prec in (("f64_sparse", pbso.PREC_F64), ("tf32x3_dense", pbso.PREC_TF32X3)):
        it = pbso.ModalIntegrator(M, synth.H, a, b); it.set_transfer(tr)
        for i in range(5):
            md.storm_buffer(it, vids[i % 16], vns[i % 16], T, prec)
        it2 = pbso.ModalIntegrator(M, synth.H, a, b); it2.set_transfer(tr)
        out = []
        lat = np.empty(n_buffers); kms = []
        for i in range(n_buffers):
            t0 = time.perf_counter()
            y, _ = md.storm_buffer(it2, vids[i % 16], vns[i % 16], T, prec)
            lat[i] = time.perf_counter() - t0
            if i < 32: out.append(y[0].copy())
            if i % 16 == 0: kms.append(md.last_kernel_ms())
        ys[name] = np.concatenate(out)
        us = lat * 1e6
        res[name] = {"p50_us": float(np.percentile(us, 50)), "p99_us": float(np.percentile(us, 99)), "impulses_per_s": float(B / np.mean(lat)),
                     "projection_kernels_ms": float(np.median(kms)), "x_realtime": float((T / synth.SAMPLE_RATE) / np.mean(lat))}
        it.close(); it2.close()
    k5 = res["tf32x3_dense"]["projection_kernels_ms"]
    res["tf32x3_dense"]["issued_tflops"] = 3 * 2.0 * M * K * B / (k5 * 1e-3) / 1e12
    res["tf32x3_dense"]["frac_of_tf32_peak"] = res["tf32x3_dense"]["issued_tflops"] / tf32_peak
    d = ys["tf32x3_dense"] - ys["f64_sparse"]
    return {"workload": "cfg3: %d modes x %d vertices, %d vertex impulses per %d-sample buffer (100 k impulses/s), host pointers in and out of pbso_modes_storm_buffer" % (M, V, B, T),
            "buffers": n_buffers, "budget_us": 1e6 * T / synth.SAMPLE_RATE, "paths": res,
            "parity_tf32x3_vs_f64_sparse": {"rel_l2": float(np.linalg.norm(d) / np.linalg.norm(ys["f64_sparse"])),
                                            "max_abs": float(np.max(np.abs(d)) / np.max(np.abs(ys["f64_sparse"]))), "buffers": 32},
            "note": "projection_kernels_ms covers the zero-fill of the dense load vectors, the scatter, the TF32 split of F, the tensor-core GEMM and the load sum"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openpbso_b200 import build
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if rank == 0:
        build.build()                 # no-op when the in-tree .so is current
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import openpbso_b200 as pbso
    from openpbso_b200 import synth
    pbso.set_device(local)
    prec = {"f32_tiled": pbso.PREC_F32_TILED, "f64": pbso.PREC_F64, "tc3x": pbso.PREC_TC3X}[args.precision]

    # the product's own multi-GPU path: pbso_comm_* (NCCL called from the C ABI on the render stream); torch.distributed
    # only carries the 128-byte id to the ranks and the barriers / max-over-ranks of the measurement
    comm = None
    if world > 1:
        uid = [pbso.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = pbso.Comm(world, rank, uid[0])
    lo, hi = comm.shard(args.objects) if comm else (0, args.objects)
    n_local = hi - lo
    # synthetic workload: generated per rank for its own shard (seeded by object range)
    w = synth.batch_workload(args.objects, args.modes, args.buffers, 1005)
    sl = slice(lo, hi)
    a, b = w["a"][sl], w["b"][sl]
    # pinned host staging for the e2e arm
    def pinned(x):
        t = torch.empty(x.shape, dtype=torch.float64 if x.dtype == np.float64 else torch.int32, pin_memory=True)
        t.numpy()[...] = x
        return t.numpy()
    space_h = pinned(w["space"][sl]); trans_h = pinned(w["trans"][sl])
    obj_h = pinned(np.arange(n_local, dtype=np.int32)); buf_h = pinned(w["imp_buf"][sl].astype(np.int32))
    n_samples = args.buffers * BUF
    br = pbso.BatchRenderer(synth.H, a, b)
    # a dedicated (non-default) stream shared by the render kernel, the NCCL reduce and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    br.set_stream(stream.cuda_stream)
    br.set_transfer(trans_h)
    br.set_impulses(obj_h, buf_h, space_h)
    mix = torch.zeros(n_samples, dtype=torch.float64, device="cuda")     # also the NCCL send buffer
    mix_host = torch.empty(n_samples, dtype=torch.float64, pin_memory=True)

    def step_device(precision=prec):
        br.render_mix_device(BUF, args.buffers, mix.data_ptr(), precision)
        if comm:
            comm.reduce_audio(mix.data_ptr(), n_samples, 0, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()          # started early: nvidia-smi needs ~0.2 s to come up
    pbso.flush_l2(256 << 20)                                # allocates the flush scratch outside the timed region
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # ---- timed region: exactly K steps, device-resident inputs -----------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        pbso.flush_l2(256 << 20)                       # evict L2 between timed iterations (default stream)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        km, launches_per_render = br.last_kernel_ms()   # syncs on the render kernels' own event pair
        kernel_ms.append(km)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    kms = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX); dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_per_step = tot.item() / args.steps
    mode_samples = float(args.objects) * args.modes * n_samples
    value = mode_samples / (ms_per_step * 1e-3)

    # ---- e2e: same job through the host-pointer C ABI, copies inside the timed region --------
    def step_e2e():
        # pinned inputs, copies enqueued on the render stream: one host synchronisation per step, after the mix is back
        br.set_transfer(trans_h, wait=False)               # H2D
        br.set_impulses(obj_h, buf_h, space_h, wait=False)  # H2D (+ host-side CSR build)
        step_device()
        if rank == 0:
            mix_host.copy_(mix, non_blocking=True)     # D2H of the result
        torch.cuda.synchronize()
    step_e2e(); barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_serial_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_serial_t, op=dist.ReduceOp.MAX)
    # The same K steps as a production renderer would run them: two handles (each with its own device input buffers, mix
    # buffer, pinned result buffer and stream) take the steps in turn, so step k+1's inputs upload over PCIe while step k
    # renders; the host waits for step k-1's result -- one synchronisation per step, every step's H2D and D2H inside the
    # timed region, results delivered one step late.  The NCCL reduces stay in issue order through an event chain.
    lanes = []
    for i in range(2):
        st_i = stream if i == 0 else torch.cuda.Stream()
        br_i = br if i == 0 else pbso.BatchRenderer(synth.H, a, b)
        if i: br_i.set_stream(st_i.cuda_stream)
        lanes.append(dict(br=br_i, st=st_i, mix=mix if i == 0 else torch.zeros_like(mix),
                          host=mix_host if i == 0 else torch.empty_like(mix_host).pin_memory(), done=torch.cuda.Event(), red=torch.cuda.Event()))
    def issue(k):
        ln, prev = lanes[k & 1], lanes[(k + 1) & 1]
        ln["br"].set_transfer(trans_h, wait=False)
        ln["br"].set_impulses(obj_h, buf_h, space_h, wait=False)
        if comm:
            # the previous step's reduce goes first: the render kernel is persistent and fills every SM (registers included), so
            # an NCCL kernel that becomes ready behind it would wait for the whole render
            ln["st"].wait_event(prev["red"])
        ln["br"].render_mix_device(BUF, args.buffers, ln["mix"].data_ptr(), prec)
        if comm:
            comm.reduce_audio(ln["mix"].data_ptr(), n_samples, 0, ln["st"].cuda_stream)
            ln["red"].record(ln["st"])
        if rank == 0:
            with torch.cuda.stream(ln["st"]):
                ln["host"].copy_(ln["mix"], non_blocking=True)
        ln["done"].record(ln["st"])
    for ln in lanes:
        ln["red"].record(ln["st"])
    issue(0); issue(1); torch.cuda.synchronize(); barrier()          # warm both lanes (second handle builds its tables here)
    t0 = time.perf_counter()
    for k in range(args.steps):
        issue(k)
        if k > 0:
            lanes[(k - 1) & 1]["done"].synchronize()                  # step k-1's mix is in host memory
    lanes[(args.steps - 1) & 1]["done"].synchronize()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_checksum = float(lanes[(args.steps - 1) & 1]["host"].abs().sum().item()) if rank == 0 else 0.0
    torch.cuda.set_stream(stream)
    e2e_value = mode_samples / (e2e_t.item() / args.steps)
    h2d = (space_h.nbytes + trans_h.nbytes + obj_h.nbytes + buf_h.nbytes)
    h2d_t = torch.tensor([float(h2d)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)

    # ---- parity of the timed configuration itself: the same job re-rendered by the FP64 direct-form kernel (the
    # reference's arithmetic: modal_integrator.h:109-110 + modal_solver.h:267-269), reduced the same way -----------
    parity = None
    if not args.no_parity and prec != pbso.PREC_F64:
        y_fast = mix.clone() if rank == 0 else None              # rank 0 holds the reduced mix of the last step
        step_device(pbso.PREC_F64)
        torch.cuda.synchronize()
        if rank == 0:
            d = (y_fast - mix)
            full = mix.abs().max().item()
            tail = slice(-44100, None)
            parity = {"vs": "k_batch_f64 (FP64 direct-form kernel, same %d objects, reduced over the same ranks)" % args.objects,
                      "rel_l2": (d.norm() / mix.norm()).item(), "max_abs": d.abs().max().item() / full,
                      "rel_l2_last_second": (d[tail].norm() / mix[tail].norm()).item(),
                      "tolerance": {"rel_l2": 1e-5, "max_abs": 1e-6}, "tc_gain_calibrated": pbso.tc_gain()}
    if rank != 0:
        if world > 1:
            dist.barrier(); comm.close(); dist.destroy_process_group()
        return
    # ---- rank 0 only: peaks, roofline, CPU baseline, real-time latency -----------------------
    checksum = float(mix_host.abs().sum().item())
    peaks, peaks_kind = measured_peaks()
    fma = {}
    for kind, name in ((0, "ffma_uniform"), (1, "ffma2_uniform"), (3, "ffma_3reg"), (4, "ffma2_3reg"), (2, "dfma")):
        t, mhz = pbso.measure_fma_peak(kind)
        fma[name] = round(t, 2)
    fp32_peak = max(fma["ffma_uniform"], fma["ffma2_uniform"], fma["ffma_3reg"], fma["ffma2_3reg"])
    info = pbso.device_info()
    k_ms = kms.item()
    ms_local = float(n_local) * args.modes * n_samples
    achieved = ms_local * FLOP_PER_MODE_SAMPLE / (k_ms * 1e-3) / 1e12          # algorithmic: 8 FLOP per mode-sample
    nominal_fp32 = info["sm_count"] * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    tf32_peak_measured, tf32_cyc, _ = pbso.measure_tc_peak(0, 2, 128)
    tf32_peak_sustained = pbso.measure_tc_peak_sustained(0, 2, 128, 300.0)
    if prec == pbso.PREC_TC3X:
        # tensor-pipe roofline: the kernel issues 3 TF32 MMAs (hi*hi, hi*lo, lo*hi) per product and 2 K columns per
        # mode (Re, Im of the tile-start state), i.e. 12 tensor FLOP per mode-sample.  The denominator is the kind::tf32
        # rate of a bare tcgen05.mma loop measured in THIS run (pbso_measure_tc_peak: one persistent CTA per SM, A in
        # TMEM, B in shared memory, nothing else running); half the cuBLAS bf16 burst figure of MEASURED_PEAKS.json is
        # reported beside it (MEASURED_PEAKS.json has no TF32 number).
        issued = ms_local * 12.0 / (k_ms * 1e-3) / 1e12
        half_bf16 = 0.5 * float(peaks.get("bf16_tflops", 1590.0))
        roofline = {"bound": "tensor", "kernel": "k_batch_tc<2> (tcgen05.mma cta_group::2 kind::tf32, 3xTF32, 12 MMAs per 16-mode K chunk and CTA pair)", "achieved": issued,
                    "peak": tf32_peak_sustained, "unit": "TFLOP/s", "frac": issued / tf32_peak_sustained,
                    "peak_burst": tf32_peak_measured, "frac_of_burst_peak": issued / tf32_peak_measured,
                    "traffic": TC3X_DRAM_BYTES_PER_MODE_SAMPLE * ms_local if (TC3X_DRAM_BYTES_PER_MODE_SAMPLE and args.modes == N_MODES and args.buffers == N_BUF) else None,
                    "kernel_ms": k_ms,
                    "peak_source": "kind::tf32 rate of a bare tcgen05.mma cta_group::2 loop measured in this run (pbso_measure_tc_peak_sustained: launched back to "
                                   "back for 300 ms, i.e. under the same power cap as the timed steps); peak_burst is one 2.6 ms launch of the same loop (%.1f cycles per 256x128x8 MMA)" % tf32_cyc,
                    "frac_of_half_bf16_cublas": issued / half_bf16, "half_bf16_cublas_tflops": half_bf16, "half_bf16_source": "0.5 x MEASURED_PEAKS.json bf16_tflops (%s, burst)" % peaks_kind,
                    "achieved_counts": "issued tensor FLOP: 3 MMAs x 2 K-columns x 2 FLOP = 12 per mode-sample",
                    "traffic_source": "dram__bytes_read+write of one ncu --set full capture of this kernel, scaled per mode-sample (profiles/r2_k_batch_tc.md)",
                    "algorithmic_tflops": achieved, "algorithmic": "8 FLOP per mode-sample (reference recurrence + dot, SURVEY 8d)",
                    "fp32_fma_peak": fp32_peak, "frac_of_fp32_fma_peak": achieved / fp32_peak, "peak_nominal_fp32": nominal_fp32,
                    "fma_microbench_tflops": fma}
    else:
        roofline = {"bound": "fp32_fma", "kernel": "k_batch_pow_g<16,16,2,ffma2>" if prec == pbso.PREC_F32_TILED else "k_batch_f64",
                    "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                    "traffic": None, "kernel_ms": k_ms,
                    "peak_source": "FMA micro-benchmark in this run (pbso_measure_fma_peak; MEASURED_PEAKS.json has no FP32 figure)",
                    "peak_nominal": nominal_fp32,
                    "fma_microbench_tflops": fma,
                    "algorithmic": "8 FLOP per mode-sample (reference recurrence); the kernel evaluates each sample from "
                                   "precomputed pole powers with 2 FMA, so frac can exceed the FMA-pipe utilisation"}
    line = {
        "metric": "mode-samples/s (IIR+FFAT)", "value": value, "unit": "mode-samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {pbso.PREC_TC3X: "tf32x3 (tcgen05 kind::tf32 hi/lo split, fp32 accumulate) + f64 carrier", pbso.PREC_F32_TILED: "f32 tiles + f64 carrier"}.get(prec, "f64"),
        "data": "synthetic",
        "config": {"workload": "cfg5 offline batch: %d objects x %d modes x %d samples (%d buffers of %d), one PointForce "
                               "per object, static listeners, mixed down" % (args.objects, args.modes, n_samples, args.buffers, BUF),
                   "parallelism": "objects block-partitioned over %d rank(s); one NCCL reduce(sum) of the FP64 mix" % world,
                   "l2": "flushed with a 256 MiB write between timed steps; per-step inputs %.0f MB" % (
                       (7 * a.size * 8 + space_h.nbytes) / 1e6),
                   "mix_abs_sum": checksum},
        "roofline": roofline, "clocks": clocks, "gpu_launches": args.steps * world * int(launches_per_render),
        "gpu_launches_note": "kernels of libpbso_b200.so inside the timed region: %d per render and rank (tc3x: k_tc_carrier, k_tc_impulse, k_batch_tc)" % int(launches_per_render),
        "e2e": {"value": e2e_value, "unit": "mode-samples/s", "h2d_bytes_per_step": int(h2d_t.item()),
                "d2h_bytes_per_step": int(mix_host.numel() * 8), "ms_per_step": 1e3 * e2e_t.item() / args.steps,
                "pipelining": "two handles take the steps in turn: step k+1's inputs upload while step k renders; one host synchronisation per step (on step k-1's result); every step's H2D and D2H inside the timed region",
                "serial_ms_per_step": 1e3 * e2e_serial_t.item() / args.steps, "serial_value": mode_samples / (e2e_serial_t.item() / args.steps),
                "mix_abs_sum_last_step": e2e_checksum},
        "wall_ms_per_step": 1e3 * t_wall / args.steps, "peaks": {"source": peaks_kind, "hbm_gbs": peaks.get("hbm_gbs")},
    }
    if parity:
        line["parity"] = parity
    if comm:
        line["config"]["parallelism"] = ("objects block-partitioned over %d ranks (pbso_comm_shard); one ncclReduce(sum) of the FP64 mix issued by "
                                         "pbso_comm_reduce_audio on the render stream (NCCL %d)" % (world, comm.nccl_version()))
    if not args.no_kernels and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_kernels
        line["kernels"] = bench_kernels.run_all(tf32_peak=tf32_peak_measured)
    if not args.no_cpu_baseline:
        model, cores = cpu_info()
        ms, dt = cpu_render_sample(cores, 1, args.modes, args.buffers, 1005)
        line["cpu_baseline"] = {"value": ms / dt, "unit": "mode-samples/s", "cores": cores, "kind": "port",
                                "sample": "%d objects x %d modes x %d samples (1 object per host thread, %s), %.1f s" % (
                                    cores, args.modes, n_samples, model, dt)}
        rh = reference_headers_entry(cores, args.modes, args.buffers)
        if rh:
            line["cpu_baseline"]["reference_headers"] = rh
    if not args.no_realtime:
        line["realtime"] = realtime_latency(pbso, synth)
        line["moving_listeners"] = moving_listeners_latency(pbso, synth)
        line["contact_storm"] = contact_storm(pbso, synth, tf32_peak_measured)
    emit(line)
    if world > 1:
        dist.barrier(); comm.close(); dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


if __name__ == "__main__":
    # Libraries write banners to stdout (NCCL prints "NCCL version ..." when the first communicator is created): keep the
    # original stdout for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
