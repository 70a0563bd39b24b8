"""Micro-benchmarks of the non-dominant kernels of the path, each against the roofline that bounds it: K5 batched
projection on tensor cores (TFLOP/s, algorithmic 2 M K B, against the kind::tf32 peak measured in the same run), K4 dense
GEMV, K3 FFAT evaluation and K6 FFAT map construction (GB/s of algorithmic bytes against the measured HBM copy bandwidth).
Every entry carries its own parity number against an FP64 evaluation done in the same run.

    python scripts/bench_kernels.py [--quick | --ffat-only | --fit-only]      prints one JSON object
    bench.py imports run_all() and folds the result into its JSON line under "kernels"."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_all(quick=False, ffat_only=False, fit_only=False, tf32_peak=None):
    import ctypes as C
    import torch
    import openpbso_b200 as pbso
    from openpbso_b200 import synth

    _sweep = torch.ones(64 << 20, dtype=torch.float32, device="cuda")      # 256 MiB read sweep

    def ev_time(fn, iters=10, warm=3, flush=True):
        s = torch.cuda.current_stream()
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            if flush:
                # evict with a 256 MiB write, then sweep a second buffer with reads so that what is left in L2 is CLEAN:
                # dirty flush lines would otherwise be written back during the timed kernel and bill it for their traffic
                pbso.flush_l2(256 << 20); _sweep.sum(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            # a ~50 us spin ahead of the first event lets the host enqueue fn()'s launches while the GPU is still busy, so the
            # event pair brackets device execution (incl. gaps between dependent kernels), not host launch latency
            torch.cuda._sleep(100000)
            e0.record(s); fn(); e1.record(s); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))

    out = {}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    out["peaks"] = {"hbm_gbs": hbm, "hbm_source": "MEASURED_PEAKS.json" if "gpu_name" in peaks else "fallback"}
    prev_stream = torch.cuda.current_stream()
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    if tf32_peak is None and not ffat_only:
        tf32_peak = pbso.measure_tc_peak(0, 1, 128)[0]
    out["peaks"]["tf32_tflops"] = tf32_peak
    ffat_only = ffat_only or fit_only
    # ---- K5 / K4: cfg3 sizes -----------------------------------------------------------------
    M, V = 2048, 20000; K = 3 * V
    if quick: M, K = 1024, 6000
    if ffat_only: M, K = 128, 64
    rng = np.random.default_rng(1003)
    U = rng.standard_normal((M, K))
    md = pbso.ModeShapes(U)
    k5 = []
    for B in ([] if ffat_only else [64, 580] if quick else [8, 64, 580, 4096]):
        F = torch.randn(B, K, device="cuda", dtype=torch.float32)
        Y = torch.empty(B, M, device="cuda", dtype=torch.float32)
        fn = lambda: md.project_dense_device(F.data_ptr(), B, Y.data_ptr(), stream_ptr=sp)
        med, best = ev_time(fn, iters=8)
        # parity on a few columns against float64 torch
        ref = (F[:4].double() @ torch.from_numpy(U).cuda().T)
        err = ((Y[:4].double() - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
        tf = 2.0 * M * K * B / (med * 1e-3) / 1e12
        k5.append({"B": B, "ms": med, "ms_best": best, "tflops": tf, "issued_tflops": 3.0 * tf, "frac_of_tf32_peak": 3.0 * tf / tf32_peak,
                   "impulses_per_s": B / (med * 1e-3), "parity_col_rel_l2_vs_fp64": err})
    out["K5_project_tc"] = {"M": M, "K": K, "bound": "tensor", "peak_tflops": tf32_peak, "runs": k5,
                            "note": "includes the per-call TF32 hi/lo split of F; algorithmic FLOP = 2 M K B, 3 MMAs issued per product: frac = 3 x algorithmic / measured kind::tf32 peak; B <= 64 is HBM-bound (U read once)"}

    f = rng.standard_normal((1, K))
    kms = []; t0 = time.perf_counter(); n = 20
    for _ in range(n):
        pbso.flush_l2(256 << 20); torch.cuda.synchronize()
        yv = md.project_dense(f); kms.append(md.last_kernel_ms())
    k4_err = float(np.linalg.norm(yv[0] - U @ f[0]) / np.linalg.norm(U @ f[0]))
    dt = (time.perf_counter() - t0) / n
    alg = (M * K * 8 + K * 8 + M * 8)
    out["K4_gemv_f64"] = {"kernel_ms": float(np.median(kms)), "algorithmic_GB": alg / 1e9, "GBps": alg / (np.median(kms) * 1e-3) / 1e9,
                          "bound": "hbm", "frac_of_hbm": alg / (np.median(kms) * 1e-3) / 1e9 / hbm, "hbm_peak_gbs": hbm, "parity_rel_l2_vs_fp64": k4_err,
                          "host_call_ms_incl_copies_and_flush": dt * 1e3}

    # ---- K4 sparse: the contact storm as vertex impulses (cfg3: 100k impulses/s = 580 per 256-sample buffer) ----------------
    if not ffat_only:
        Vn = K // 3
        k4s = []
        for B in ([580] if quick else [1, 580, 4096]):
            vids = rng.integers(0, Vn, B).astype(np.int32); vns = synth.unit_vectors(B, 1003)
            md.project_vertices(M, vids, vns)
            ts = []
            for _ in range(10):
                t0 = time.perf_counter(); md.project_vertices(M, vids, vns); ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            k4s.append({"B": B, "host_call_us": dt * 1e6, "impulses_per_s": B / dt, "d2h_MB": B * M * 8 / 1e6,
                        "note": "host pointers in and out: B x 2048 modal loads (doubles) returned per call"})
        out["K4_project_sparse"] = {"M": M, "V": Vn, "runs": k4s}

    # ---- K3: cfg4 (1024 modes x 64 listeners) and the HUD sphere (10242 listeners) ---------------
    Mf = 1024
    freqs = synth.mode_frequencies(Mf, 1004)
    fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(freqs, 2000))
    k3 = []
    for L in ([] if fit_only else [10242] if ffat_only else [64] if quick else [1, 64, 10242]):
        pos = torch.from_numpy(synth.listeners(L, 5)).cuda()
        o = torch.empty(L, Mf, device="cuda", dtype=torch.float64)
        fn = lambda: pbso._capi.check(pbso.lib().pbso_ffat_eval_device(fm._h, Mf, C.c_void_p(pos.data_ptr()), L, C.c_void_p(o.data_ptr()), C.c_void_p(sp)))
        med, best = ev_time(fn, iters=10)
        D = 6144
        bytes_alg = Mf * (min(D, 4 * L) * 8 + L * 8)
        k3.append({"L": L, "us": med * 1e3, "us_best": best * 1e3, "algorithmic_MB": bytes_alg / 1e6, "GBps": bytes_alg / (med * 1e-3) / 1e9, "frac_of_hbm": bytes_alg / (med * 1e-3) / 1e9 / hbm,
                   "kernel": "k_ffat_locate + k_ffat_tiles" if L >= 2048 else "k_ffat_gather_fused" if L <= 256 else "k_ffat_locate + k_ffat_gather"})
        if L >= 2048:
            o_tiles = o.clone()
            os.environ["PBSO_FFAT_GATHER"] = "1"
            med_g, _ = ev_time(fn, iters=10)
            del os.environ["PBSO_FFAT_GATHER"]
            k3[-1]["per_listener_gather_us"] = med_g * 1e3
            k3[-1]["parity_max_rel_vs_per_listener_gather"] = float(((o_tiles - o).abs() / o.abs().clamp_min(1e-300)).max().item())
        else:
            k3[-1]["note"] = "launch / latency bound at this size: %.1f MB in one launch" % (bytes_alg / 1e6)
    # the 8-bit view (FFAT_Map::Compress): one byte per texel + maxAmp/255 per (map, face)
    k3q = []
    if not fit_only:
        dicts = synth.ffat_maps(freqs, 2000)
        fm8 = pbso.FFATMaps.from_dicts(dicts); fm8.Compress()
        fd8 = pbso.FFATMaps.from_dicts([dict(m, psi=fm8.get_compressed(i)[2], is_compressed=True) for i, m in enumerate(dicts)])
        for L in ([10242] if ffat_only else [64] if quick else [64, 10242]):
            pos = torch.from_numpy(synth.listeners(L, 5)).cuda()
            o = torch.empty(L, Mf, device="cuda", dtype=torch.float64); o2 = torch.empty_like(o)
            fn = lambda: pbso._capi.check(pbso.lib().pbso_ffat_eval_device_view(fm8._h, Mf, C.c_void_p(pos.data_ptr()), L, 1, C.c_void_p(o.data_ptr()), C.c_void_p(sp)))
            med, best = ev_time(fn, iters=10)
            os.environ["PBSO_FFAT_GATHER"] = "1"
            pbso._capi.check(pbso.lib().pbso_ffat_eval_device_view(fd8._h, Mf, C.c_void_p(pos.data_ptr()), L, 1, C.c_void_p(o2.data_ptr()), C.c_void_p(sp)))
            del os.environ["PBSO_FFAT_GATHER"]
            torch.cuda.synchronize()
            bytes_alg = Mf * (min(6144, 4 * L) * 1 + 6 * 8 + L * 8)
            k3q.append({"L": L, "us": med * 1e3, "us_best": best * 1e3, "algorithmic_MB": bytes_alg / 1e6, "GBps": bytes_alg / (med * 1e-3) / 1e9,
                        "frac_of_hbm": bytes_alg / (med * 1e-3) / 1e9 / hbm, "kernel": "k_ffat_gather_fused<u8>" if L <= 256 else "k_ffat_locate + k_ffat_gather_q8x4",
                        "parity_max_rel_vs_the_doubles_of_compressed_Psi": float(((o - o2).abs() / o2.abs().clamp_min(1e-300)).max().item())})
    out["K3_ffat_eval"] = {"modes": Mf, "texels": 6144, "bound": "hbm", "hbm_peak_gbs": hbm, "runs": k3,
                           "algorithmic_bytes": "M (min(D, 4 L) 8 + L 8) per call (SURVEY 8d)",
                           "compressed_view_u8": {"runs": k3q, "algorithmic_bytes": "M (min(D, 4 L) 1 + 48 + L 8) per call: one byte per texel + 6 face scales + the output"}}

    # ---- K6: FFAT map construction (FFAT_Map<T,3>::Solve for all modes of an object at once) -------------
    # shells of 16/24/32 cells per edge (shell 2 = the 6 x 32 x 32 run-time map of the other configs), 1024 modes
    nm = 64 if quick else 1024
    w = synth.ffat_fit_workload(nm, 1006)
    ft = pbso.FFATFitter(w["cell_size"], w["V"], w["n_elements"])
    dk = torch.from_numpy(w["k"]).cuda()
    dp = torch.from_numpy(np.ascontiguousarray(w["pressure"]).view(np.float64)).cuda()
    dpsi = torch.empty(nm, ft.n_directions, dtype=torch.float64, device="cuda")
    dsc = torch.empty(nm, dtype=torch.float64, device="cuda")
    k6 = []
    dp_packed = torch.from_numpy(np.ascontiguousarray(w["pressure"][:, 0::2]).view(np.float64)).cuda()   # one complex per quad
    ref_psi = {}
    for layout, scaling, defer in (("reference", False, False), ("reference", True, False), ("packed", False, False), ("packed", True, False), ("packed", True, True)):
        packed = layout == "packed"
        fn = lambda: ft.solve_device(nm, dk.data_ptr(), (dp_packed if packed else dp).data_ptr(), dpsi.data_ptr(), scaling, dsc.data_ptr(), sp,
                                     packed=packed, defer_scale=defer)
        med, best = ev_time(fn, iters=10)
        out_psi = dpsi * dsc[:, None] if defer else dpsi.clone()
        if not packed: ref_psi[scaling] = out_psi
        alg = nm * (16 * ft.n_elements_total + 8 * ft.n_directions)          # complex samples read + Psi written
        touched = nm * ((16 if packed else 32) * ft.n_elements_total + 8 * ft.n_directions)   # the reference layout interleaves unused entries
        k6.append({"layout": layout, "power_scaling": scaling, "deferred_scale": defer, "us": med * 1e3, "algorithmic_MB": alg / 1e6, "GBps": alg / (med * 1e-3) / 1e9,
                   "frac_of_hbm": alg / (med * 1e-3) / 1e9 / hbm, "sector_MB": touched / 1e6,
                   "sector_GBps": touched / (med * 1e-3) / 1e9, "sector_frac_of_hbm": touched / (med * 1e-3) / 1e9 / hbm,
                   "parity_max_rel_vs_reference_layout": float((out_psi / ref_psi[scaling] - 1.0).abs().max().item())})
    dpsi_keep = ref_psi[True]
    t0 = time.perf_counter(); psi_h, _ = ft.Solve(w["k"], w["pressure"], True); host_ms = (time.perf_counter() - t0) * 1e3
    k6_err = float(np.max(np.abs(dpsi_keep.cpu().numpy() / psi_h - 1.0)))
    out["K6_ffat_fit"] = {"bound": "hbm", "hbm_peak_gbs": hbm, "parity_max_rel_device_entry_vs_host_entry": k6_err, "modes": nm, "shells": ft.n_shells, "n_elements_total": ft.n_elements_total, "n_directions": ft.n_directions,
                          "runs": k6, "host_call_ms_incl_copies": host_ms, "host_call_kernel_ms": ft.last_kernel_ms(),
                          "note": "algorithmic bytes = 16 B per shell sample + 8 B per Psi value; sector bytes count the unused odd entries of the reference's vector layout that share a 32 B sector with each sample; layout packed = one complex per quad (PBSO_FIT_PACKED), deferred_scale = the factor is returned and Psi left unscaled (PBSO_FIT_DEFER_SCALE)"}
    torch.cuda.synchronize()
    torch.cuda.set_stream(prev_stream)
    return out



if __name__ == "__main__":
    print(json.dumps(run_all(quick="--quick" in sys.argv, ffat_only="--ffat-only" in sys.argv, fit_only="--fit-only" in sys.argv)))
