"""Single-GPU bench lines for the other BASELINE.json configs (C2, C3, C5).

`bench.py --workload c2|c3|c5` lands here.  Same JSON shape as the default (C4) line; these are the
numbers DESIGN.md quotes for variant selection and for the frame-to-frame / local-window workloads.
Inputs here fit in L2, so L2 is flushed (a 256 MB buffer is overwritten) between timed iterations
and every iteration is timed with its own CUDA-event pair.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def _events(torch, n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def measure_orb(steps=30, n_features=2000):
    """Auxiliary record for the row in front of the path: one 752 x 480 frame, keypoints from
    cv2's detector, descriptors through hm_frame_put_orb (image + keypoints H2D inside the timed region) next to
    cv2.ORB.compute on the host cores.  Not a pairs/s number: ms per frame, and the result is compared with cv2's."""
    import cv2
    import torch
    import slam_experiments_b200 as sx
    from slam_experiments_b200 import synth
    from slam_experiments_b200.feature_detectors import keypoint_arrays
    from oracle import orb_oracle
    img = synth.textured_image(480, 752, 11)
    orb = cv2.ORB.create(nfeatures=n_features)
    kps = orb.detect(img, None)
    _, ref = orb.compute(img, kps)
    xy, ang, octv = keypoint_arrays(kps)
    store = sx.FrameDescriptorStore()
    got = store.put_image("f", img, kps, want_descriptors=True)
    verified = bool(np.array_equal(got, ref)) and bool(np.array_equal(got[::16], orb_oracle.describe(img, xy[::16], ang[::16], octv[::16])))
    if not verified:
        raise SystemExit("bench: device ORB descriptors differ from cv2 / the oracle")
    ctx = store._ctx

    def put():
        ctx.frame_put_orb(0, img, xy, ang, octv, 8, None, False)
    for _ in range(5):
        put()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        put()
        torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) / steps * 1e3
    t0 = time.perf_counter()
    for _ in range(steps):
        orb.compute(img, kps)
    cpu_ms = (time.perf_counter() - t0) / steps * 1e3
    t0 = time.perf_counter()
    for _ in range(max(3, steps // 3)):
        orb.detect(img, None)
    det_ms = (time.perf_counter() - t0) / max(3, steps // 3) * 1e3
    return {"workload": "orb_describe_752x480", "keypoints": len(kps), "pyramid_levels": 8,
            "device_ms_per_frame": dev_ms, "api": "hm_frame_put_orb: host image + cv2 keypoints -> descriptors in a resident frame slot",
            "h2d_bytes_per_frame": int(img.size + len(kps) * 20), "d2h_bytes_per_frame": 0, "gpu_launches_per_frame": 10,
            "cpu_baseline": {"compute_ms_per_frame": cpu_ms, "detect_ms_per_frame": det_ms, "engine": f"cv2.ORB {cv2.__version__}",
                             "cores": cv2.getNumThreads(), "kind": "reference",
                             "note": "detection stays on the host in both arms; compute is the stage the device replaces"},
            "verified_vs_oracle": verified, "verified": "all descriptors bit-identical to cv2.ORB.compute; every 16th against the oracle"}


def run(args):
    line = measure(args.workload, args.steps, args.warmup, args.variant, args.n)
    print(json.dumps(line), flush=True)
    return 0


def measure(wl, steps, warmup, variant_arg="auto", n_arg=65536, cpu_seconds=10.0):
    """One bench record (the dict bench.py prints) for workload c1 | c2 | c3 | c5 on cuda:0."""
    import types
    args = types.SimpleNamespace(workload=wl, steps=steps, warmup=warmup, variant=variant_arg, n=n_arg)
    import torch
    import cv2
    import slam_experiments_b200 as sx
    from slam_experiments_b200 import _native as nat, synth
    from bench import ClockSampler, measured_peaks, tensor_roofline, METRIC, UNIT, POPC_PER_PAIR
    from oracle import c_oracle

    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    peaks, peak_src = measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps, warm = args.steps, max(args.warmup, 3)
    verified_what = ""

    if wl == "c3":
        n = args.n
        q, t = synth.sweep(n, "U")
        qd, td = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
        pairs = float(n) * n
        variant = args.variant if args.variant != "auto" else nat.select_variant(n, n, 1)
        fn = lambda: nat.knn2_keys(qd, td, variant=variant)
        bf = sx.BFMatcher(cv2.NORM_HAMMING, variant=args.variant)
        e2e_fn = lambda: bf.knnMatch(q, t, k=2)
        e2e_arr = lambda: bf.knn_tensors(q, t, 2)
        h2d, d2h = 2 * n * 32, n * 16
        cfg = {"workload": f"c3_sweep_{n}x{n}", "distribution": "uniform", "variant": variant}
        rows_chk = np.linspace(0, n - 1, min(n, 1024)).astype(np.int64)
        check = lambda: np.array_equal(fn().cpu().numpy().view(np.uint64)[rows_chk], c_oracle.knn2_keys(q[rows_chk], t))
        verified_what = f"{rows_chk.size} of {n} rows vs the C oracle"
        cpu_fn = lambda m: cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q[:m], t, k=2)
        cpu_pairs = lambda m: float(m) * n
        launches = 2 if variant == "f4" else 3 if variant == "i8" else 2      # f4: train expansion + one k-NN launch (query expanded inside, splits merged inside)
        nq_k, nt_k = n, n
    elif wl == "c1":
        # BASELINE configs[0]: the reference's own bundled pair (1.png / 2.png, ORB 200 features as shipped in
        # main.py:35).  The descriptors and the reference's outputs come from the committed golden fixture
        # (tests/golden/c1_orb200.npz, generated from /root/reference by tests/golden/make_golden.py).
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_orb200.npz"))
        q, t = np.ascontiguousarray(g["query"]), np.ascontiguousarray(g["train"])
        qd, td = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
        nb = 1
        pairs = float(q.shape[0]) * t.shape[0]
        variant = args.variant if args.variant != "auto" else nat.select_variant(q.shape[0], t.shape[0], 1)
        fn = lambda: nat.knn2_keys(qd, td, variant=variant)
        m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, variant=args.variant)
        e2e_fn = lambda: m.match(t, q)                       # feature_matchers.py:36-44: match(source = train, query)
        e2e_arr = lambda: m.match_tensors(t, q)
        h2d, d2h = (q.shape[0] + t.shape[0]) * 32, q.shape[0] * 12
        cfg = {"workload": "c1_bundled_pair_orb200", "rows": [int(q.shape[0]), int(t.shape[0])], "variant": variant,
               "note": "latency-bound: 40,000 pairs; report the latency, not a roofline fraction"}

        def check():
            out = m.match(t, q)
            rows = [(x.queryIdx, x.trainIdx, x.imgIdx, int(x.distance)) for x in out]
            return rows == [tuple(r) for r in g["ref_match"].tolist()]
        verified_what = "all 200 matches vs the reference's recorded output (tests/golden/c1_orb200.npz)"
        ref = cv2.BFMatcher(cv2.NORM_HAMMING)
        cpu_fn = lambda k: [ref.match(q, t) for _ in range(k)]
        cpu_pairs = lambda k: float(k) * pairs
        launches = 2
        nq_k, nt_k = q.shape[0], t.shape[0]
    elif wl == "c2":
        # 100 warped 752 x 480 frames, real ORB (tests/golden/make_golden_orb.py generated the
        # descriptors from the reference's detector; bit density 0.54, correlated rows)
        frames = synth.euroc_shaped_sequence()
        fd = torch.from_numpy(frames).to(dev)
        nb = frames.shape[0] - 1
        pairs = float(nb) * 2000 * 2000
        variant = args.variant if args.variant != "auto" else nat.select_variant(2000, 2000, nb)
        fn = lambda: nat.knn2_keys_batched(fd[1:], fd[:-1], variant=variant)
        m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, variant=args.variant)

        def e2e_fn():
            out = 0
            for i in range(nb):                      # frontend.py:181-187, one call per new frame
                out += len(m.match(frames[i], frames[i + 1]))
            return out

        def e2e_arr():
            for i in range(nb):
                m.match_tensors(frames[i], frames[i + 1])

        # 8f ranks 1 + 2: every frame uploaded once (descriptors + keypoint positions), the matched point
        # arrays of utils.py:13-19 gathered on the device; only the new frame crosses PCIe per step
        positions = np.random.default_rng(7).integers(0, 752, (frames.shape[0], 2000, 2)).astype(np.int32)
        store = sx.FrameDescriptorStore(capacity=4, variant=args.variant)

        def e2e_points():
            store.put(0, frames[0], positions[0])
            for i in range(nb):
                store.put(i + 1, frames[i + 1], positions[i + 1])
                store.matched_points(i, i + 1)
        extra_e2e = {"resident_store_points_out": e2e_points}
        h2d, d2h = nb * 2 * 2000 * 32, nb * 2000 * 12
        cfg = {"workload": "c2_euroc_shaped_sequence", "frames": 100, "rows_per_frame": 2000, "variant": variant,
               "batched": "99 (last, current) problems in one launch over overlapping windows of the resident sequence"}
        def check():
            keys = fn().cpu().numpy().view(np.uint64)
            return all(np.array_equal(keys[i], c_oracle.knn2_keys(frames[i + 1], frames[i])) for i in range(nb))
        verified_what = "all 99 problems, all rows vs the C oracle"
        ref = cv2.BFMatcher(cv2.NORM_HAMMING)
        cpu_fn = lambda k: [ref.match(frames[i + 1], frames[i]) for i in range(k)]
        cpu_pairs = lambda k: float(k) * 2000 * 2000
        launches = 2 if variant == "f4" else 3 if variant == "i8" else 2      # f4: train expansion + one k-NN launch (query expanded inside, splits merged inside)
        nq_k, nt_k = 2000, 2000
    else:  # c5
        qs, ts = synth.local_window(32, 10000)
        qd, td = torch.from_numpy(qs).to(dev), torch.from_numpy(ts).to(dev)
        nb = 32
        pairs = float(nb) * 10000 * 10000
        variant = args.variant if args.variant != "auto" else nat.select_variant(10000, 10000, nb)
        fn = lambda: nat.match_fused(qd, td, ratio=0.75, cross_check=True, variant=variant)
        m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, ratio=0.75, cross_check=True, variant=args.variant)

        def e2e_fn():
            return sum(len(m.match(ts[i], qs[i])) for i in range(nb))

        def e2e_arr():
            for i in range(nb):
                m.match_tensors(ts[i], qs[i])
        # the window resident on the device (KeyframeWindow): only the queries cross PCIe, one batched call, one D2H
        win = sx.KeyframeWindow(nb, ratio=0.75, cross_check=True, variant=args.variant)
        for i in range(nb):
            win.put(i, ts[i])
        got = win.match_tensors(qs)
        eq0 = c_oracle.pipeline(qs[5], ts[5], 0.75, True)
        if not all(np.array_equal(a, b) for a, b in zip(got[5], eq0)):
            raise SystemExit("bench: KeyframeWindow differs from the oracle")
        extra_e2e = {"resident_window_arrays_out": lambda: win.match_tensors(qs),
                     "resident_window_dmatch_out": lambda: win.match(qs),
                     "resident_window_one_query_arrays_out": lambda: win.match_tensors(qs[0])}
        h2d, d2h = nb * 2 * 10000 * 32, nb * 10000 * 12
        cfg = {"workload": "c5_local_window_32x10k", "pipeline": "knn2 + ratio 0.75 (in the k-NN kernel) + mutual cross-check over the candidate train rows", "variant": variant,
               "pairs_counted": "Nq*Nt*batch (the candidate pass of the mutual check counts no extra pairs)"}

        def check():
            oq, ot, od, cnt = (x.cpu().numpy() for x in fn())
            for i in range(nb):
                eq, et, ed = c_oracle.pipeline(qs[i], ts[i], 0.75, True)
                k = int(cnt[i])
                if not (k == len(eq) and np.array_equal(oq[i, :k], eq) and np.array_equal(ot[i, :k], et) and np.array_equal(od[i, :k], ed)):
                    return False
            return True
        verified_what = "all 32 problems: q, t and distance of every surviving match vs the C oracle"

        def cpu_fn(k):
            from oracle import cv2_ref
            return [cv2_ref.pipeline(qs[i], ts[i], 0.75, True) for i in range(k)]
        cpu_pairs = lambda k: float(k) * 10000 * 10000
        launches = 5 if variant == "f4" else 6 if variant == "i8" else 3     # f4: prepare, forward k-NN (ratio test + candidate selection inside), prepare, candidate pass, filter
        nq_k, nt_k = 10000, 10000

    verified = bool(check())
    if not verified:
        raise SystemExit("bench: GPU result differs from the oracle")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ev = _events(torch, steps)
    kev = _events(torch, steps)
    for i in range(steps):
        flush.zero_()                                 # L2 flush between timed iterations
        nat.profile_events(*kev[i])
        ev[i][0].record()
        fn()
        ev[i][1].record()
    torch.cuda.synchronize()
    nat.profile_events(None, None)
    clocks = sampler.stop()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    kms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    def wall(f, reps):
        for _ in range(1 if wl != "c1" else 300):     # tiny calls: let host and device clocks settle first
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    reps = max(2, min(steps, 5)) if wl != "c1" else 2000
    e2e_s, e2e_arr_s = wall(e2e_fn, reps), wall(e2e_arr, reps)
    extra_s = {k: wall(f, reps) for k, f in (extra_e2e if wl in ("c2", "c5") else {}).items()}

    # cpu baseline: bounded sample, about `cpu_seconds`
    unit_probe = 8 if wl == "c3" else 1
    t0 = time.perf_counter(); cpu_fn(unit_probe); dt = max(time.perf_counter() - t0, 1e-4)
    full_units = {"c3": args.n, "c2": 99, "c5": 32, "c1": 20000}[wl]
    k = int(max(unit_probe, min(full_units, unit_probe * cpu_seconds / dt)))
    t0 = time.perf_counter(); cpu_fn(k); dt = time.perf_counter() - t0
    cpu = {"value": cpu_pairs(k) / dt / 1e9, "unit": UNIT, "cores": cv2.getNumThreads(), "kind": "reference",
           "sample": f"{k} of {full_units} {'query rows' if wl == 'c3' else 'repetitions' if wl == 'c1' else 'problems'} through cv2.BFMatcher",
           "engine": f"cv2.BFMatcher {cv2.__version__}", "seconds": dt}

    # the event pair brackets the FIRST dominant-kernel launch of the call (the forward k-NN for c5)
    kpairs = float(nq_k) * nt_k * (1 if wl == "c3" else nb)
    if variant in ("i8", "f4"):
        roof = tensor_roofline(nat, peaks, peak_src, variant, nq_k, nt_k, kms, 1 if wl in ("c3", "c1") else nb)
    else:
        ach = kpairs * POPC_PER_PAIR / (kms * 1e-3) / 1e12
        peak = nat.sm_count() * 16 * peaks["sm_max_mhz"] * 1e6 / 1e12
        roof = {"bound": "popc", "kernel": "hm_popc_knn2_kernel", "achieved": ach, "peak": peak, "unit": "Tpopc/s",
                "frac": ach / peak, "traffic": None, "kernel_ms": kms,
                "note": "POPC issue roofline: 8 POPC per pair; peak = SMs x 16/clk (measured 15.7, tools/microbench) x sm_max_mhz"}
    line = {"metric": METRIC, "value": pairs / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": dict(cfg, l2="flushed between timed iterations (256 MB overwrite)"),
            "clocks": clocks,
            "e2e": {"value": pairs / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": "drop-in match()/knnMatch(): numpy in, DMatch out",
                    "arrays_out_ms_per_step": e2e_arr_s * 1e3},
            "gpu_launches": steps * launches, "roofline": roof, "cpu_baseline": cpu, "verified_vs_oracle": verified,
            "verified": verified_what}
    if wl in ("c1", "c2", "c5"):
        line["frame_pairs_per_s"] = {"device": nb / (ms * 1e-3), "e2e_dmatch": nb / e2e_s, "e2e_arrays": nb / e2e_arr_s,
                                     "cpu": k / dt}
        for name, sec in extra_s.items():
            line["frame_pairs_per_s"]["e2e_" + name] = nb / sec
    if wl == "c5":      # the same metric through the resident window: train frames stay in HBM, only the queries cross PCIe
        line["e2e"]["resident_window"] = {
            name: {"value": pairs / sec / 1e9, "unit": UNIT, "ms_per_step": sec * 1e3,
                   "h2d_bytes_per_step": (10000 * 32) * (1 if "one_query" in name else nb), "d2h_bytes_per_step": nb * 4 + nb * 10000 * 12}
            for name, sec in extra_s.items()}
        line["e2e"]["resident_window"]["api"] = "KeyframeWindow.match_tensors / match: one batched pipeline call over the resident keyframes"
    return line
