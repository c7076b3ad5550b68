#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MAVMAP hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: the oracle restatement of the
                                                            reference's Ceres/OpenCV path, all host cores)

Primary metric (BASELINE.json): bundle-adjustment LM iterations per second on configs[1]
(500-image PINHOLE sequence, ~1 M observations).  A "step" is one LM iteration = residual+Jacobian
(K1) + Schur assembly (K2) + PCG solve (K3) + back-substitution / candidate cost (K4) + accept/reject.
The same JSON line carries the matching throughput (image pairs/s, 5k x 5k SURF-64) as `secondary`.

value    : inputs resident in HBM (mm_ba_session_*), CUDA-event timed on the session's stream
e2e      : same metric through mm_ba_solve with HOST buffers (upload + structure setup + K
           iterations + download inside the timed region)
N > 1    : BA does not need to shard (north_star: single-GPU unless HBM overflows) -> N independent
           replicas ("replicas only", weak scaling); matching shards by pair with one gather.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ba_lm_iterations_per_sec"
UNIT = "LM iterations/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        """samples taken between t0 and t1 (the warm-up + timed region); the nearest ones if the window caught none"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = [r[1:] for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1 + 0.05)]
        if not rows and self.rows:
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
            rows = [r[1:] for r in sorted(self.rows, key=lambda r: abs(r[0] - mid))[:3]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ba_config(workload):
    from mavmap_b200 import synthetic
    name = {"ba_cfg2": "cfg2", "ba_cfg4": "cfg4", "ba_cfg1": "cfg1", "ba_small": "small"}[workload]
    return name, dict(synthetic.BA_CONFIGS[name])


def options(iters, oracle=False):
    from mavmap_b200.ba import default_c_options
    if oracle:
        from oracle import orc
        o = orc.default_options()
    else:
        o = default_c_options()
    o.max_num_iterations = iters; o.function_tolerance = 0.0; o.gradient_tolerance = 0.0     # fixed iteration count
    return o


def cpu_ba_iterations(flat, iters):
    """Oracle (restated Ceres-semantics LM + Schur + Cholesky, OpenMP) timed on the host cores."""
    from oracle import orc
    o = options(iters, oracle=True)
    f = flat.copy()
    t = time.perf_counter()
    s = orc.solve_flat(f, o)
    dt = time.perf_counter() - t
    n = s.num_successful_steps + s.num_unsuccessful_steps
    return n / dt, dt, n, orc.num_threads(), s


def cpu_match_pairs(desc, n_pairs):
    import cv2
    from oracle import orc
    cv2.setNumThreads(os.cpu_count() or 1)
    t = time.perf_counter()
    for p in range(n_pairs):
        orc.match_pair_cv2(desc[p % len(desc)], desc[(p + 1) % len(desc)], ratio_test=True, max_ratio=0.9)
    return n_pairs / (time.perf_counter() - t)


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  Real Ceres/OpenCV-C++ cannot be built here (not in the image),
    so this is the oracle port of the reference path (BA) and cv2.BFMatcher (matching)."""
    if rank != 0:
        return
    from mavmap_b200 import synthetic
    name, kw = ba_config(args.workload)
    flat, _ = synthetic.make_ba_problem(**kw)
    steps = max(1, min(args.steps, args.ref_max_steps))
    ips, dt, n, cores, s = cpu_ba_iterations(flat, steps)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": 0,
            "ms_per_step": 1e3 * dt / max(n, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "%s: %d images, %d points, %d observations, PINHOLE, fixed intrinsics" % (name, flat.n_img, flat.n_pt, flat.n_obs)},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "full %s problem, %d LM iterations of the oracle (restated Ceres-semantics LM + Schur + skyline Cholesky, OpenMP; NOT Ceres)" % (name, n)},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ba_cfg2", choices=["ba_cfg2", "ba_cfg4", "ba_cfg1", "ba_small"])
    ap.add_argument("--match-pairs", type=int, default=120)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--ref-max-steps", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the extra 5000-image / 10 M observation measurement (N = 1 only)")
    ap.add_argument("--ba-mode", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: independent problems per GPU (default, north_star: BA stays on one GPU) or ONE problem with its points sharded")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the JSON line only
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda"); dist.all_reduce(warm); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    from mavmap_b200 import _lib, synthetic
    from mavmap_b200.ba import BASession, solve_flat
    from mavmap_b200.matching import MatchSet
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ BA, resident
    name, kw = ba_config(args.workload)
    sharded = args.ba_mode == "sharded" and world > 1
    if not sharded:
        kw["seed"] = kw["seed"] + 1000 * rank             # replicas: one independent problem per rank
    flat, _ = synthetic.make_ba_problem(**kw)
    W, K = args.warmup, args.steps
    stream = torch.cuda.current_stream().cuda_stream
    if sharded:
        from mavmap_b200.parallel import make_allreduce_callback
        ar_cb = make_allreduce_callback()
        sess = BASession(flat.copy(), options(W + K), stream=stream, rank=rank, world=world, allreduce=ar_cb)
    else:
        sess = BASession(flat.copy(), options(W + K), stream=stream)
    n_blocks = sess.num_blocks()
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.25)                                       # nvidia-smi needs a moment to produce its first sample
    t_load0 = time.time()
    sess.iterate(W)
    barrier()
    l0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = sess.iterate(K)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.kernel_launch_count() - l0
    clocks = sampler.stop(t_load0, time.time())
    summ = sess.summary().as_dict()
    assert done == K, "LM stopped early (%d of %d iterations)" % (done, K)
    value = (1 if sharded else world) * K / (ms * 1e-3)

    # per-kernel timings on the resident state (live, CUDA events inside the library)
    k1_ms = sess.time_kernel(0, 20); k2_ms = sess.time_kernel(1, 10); k4_ms = sess.time_kernel(2, 20); spmv_ms = sess.time_kernel(3, 50)
    coarse_dim = sess.coarse_dim()
    coarse_ms = sess.time_kernel(4, 10) if coarse_dim else 0.0
    n_obs, n_pt, n_img = flat.n_obs, flat.n_pt, flat.n_img
    k1_bytes = 184.0 * n_obs + 48.0 * n_img + 24.0 * n_pt            # K1 algorithmic bytes
    k2_bytes = 168.0 * n_obs + 72.0 * n_pt + 216.0 * n_img + 288.0 * n_blocks
    k3_bytes = 288.0 * (2 * n_blocks - n_img) + 5 * 48.0 * n_img      # both triangles are read through the CSR
    pcg_iters = sum(summ["trace_linear_iterations"][W + 1:W + 1 + K])
    lin_ms = summ["ms"]
    traffic = None                       # dram bytes per launch of K1 from the committed `ncu --set full` capture of this workload
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(name, {}).get("k_residual_jacobian<1,0>")
    roof_k1 = {"kernel": "k_residual_jacobian (K1)", "bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "traffic": traffic, "ms": k1_ms, "peak_source": peaks["source"], "algorithmic_bytes": k1_bytes}
    roof_k1["frac"] = roof_k1["achieved"] / roof_k1["peak"]
    roof_k2 = {"kernel": "k_schur_point + k_schur_blocks + k_schur_cam (K2)", "bound": "hbm", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "traffic": None, "ms": k2_ms}
    roof_k2["frac"] = roof_k2["achieved"] / roof_k2["peak"]
    roof_k3 = {"kernel": "k_pcg_spmv (K3, one PCG iteration's SpMV)", "bound": "hbm", "achieved": k3_bytes / (spmv_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "traffic": None, "ms": spmv_ms}
    roof_k3["frac"] = roof_k3["achieved"] / roof_k3["peak"]
    # the persistent solver as it runs inside the step: bytes one PCG iteration touches (blocks through both row references, five
    # vectors, the dense coarse inverse) over the measured time per iteration; S and the inverse are L2 / shared-memory resident,
    # so this is a bandwidth figure for orientation, not an HBM-bound kernel
    try:
        per_it_ms = lin_ms["pcg"] / max(int(pcg_iters), 1)
        k3_it_bytes = k3_bytes + 8.0 * coarse_dim * coarse_dim
        roof_k3p = {"kernel": "persistent PCG kernel, one iteration (K3; dominant by time)", "bound": "hbm", "achieved": k3_it_bytes / (per_it_ms * 1e-3) / 1e9 if per_it_ms > 0 else 0.0,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None, "ms": per_it_ms, "note": "operands resident in L2 / shared memory: latency- and barrier-bound (3 grid barriers per iteration)"}
        roof_k3p["frac"] = roof_k3p["achieved"] / roof_k3p["peak"]
    except Exception as e:      # orientation figure only: never let it take the headline line down
        roof_k3p = {"kernel": "persistent PCG kernel", "error": repr(e)}
    sess.close()

    # ------------------------------------------------------------------ BA, end to end through mm_ba_solve (host buffers)
    if sharded:
        kw2 = dict(kw); kw2["seed"] = kw["seed"] + 1000 * rank
        flat, _ = synthetic.make_ba_problem(**kw2)          # the host-buffer call is per GPU: one problem each
        n_obs, n_pt, n_img = flat.n_obs, flat.n_pt, flat.n_img
    warm = flat.copy(); solve_flat(warm, options(1))          # load kernels / allocator warm-up, untimed
    barrier()
    f2 = flat.copy()
    t0 = time.perf_counter()
    s2 = solve_flat(f2, options(K))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    n2 = s2.num_successful_steps + s2.num_unsuccessful_steps
    h2d = 24 * n_obs + 48 * n_img + 24 * n_pt + 72 + 6 * n_img + n_pt
    d2h = 48 * n_img + 24 * n_pt + 72
    e2e = {"value": world * n2 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(n2, 1)), "d2h_bytes_per_step": int(d2h / max(n2, 1)),
           "total_ms": 1e3 * e2e_s, "setup_ms": s2.ms_setup, "steps": n2,
           "note": "mm_ba_solve on host arrays: upload + device structure setup + %d LM iterations + download" % n2}

    # ------------------------------------------------------------------ matching (secondary), sharded by pair
    secondary = None
    if not args.no_match:
        n_feat, kdim, n_imgs = 5000, 64, 6
        desc, xy = synthetic.make_descriptors(n_imgs, n_feat, kdim, seed=0xF00D + 3)
        ms_set = MatchSet(desc, None)
        pairs_all = [(i, j) for i in range(n_imgs) for j in range(i + 1, n_imgs)]
        pairs_all = (pairs_all * (1 + args.match_pairs * world // len(pairs_all)))[: args.match_pairs * world]
        mine = pairs_all[rank::world]
        np_ = len(mine)
        cnt = torch.zeros(np_, dtype=torch.int32, device="cuda"); qd = torch.empty(np_ * n_feat, dtype=torch.int32, device="cuda")
        td = torch.empty_like(qd); dd = torch.empty(np_ * n_feat, dtype=torch.float32, device="cuda")
        def run():
            ms_set.match_pairs_device(mine, cnt.data_ptr(), qd.data_ptr(), td.data_ptr(), dd.data_ptr(), n_feat, stream, True, 0.9, -1)
            if world > 1:        # the one exchange step of the sharded path: gather counts + packed match lists over NCCL/NVLink
                dist.all_gather_into_tensor(g_cnt, cnt)
                dist.all_gather_into_tensor(g_q, qd); dist.all_gather_into_tensor(g_t, td); dist.all_gather_into_tensor(g_d, dd)
        if world > 1:
            g_cnt = torch.empty(world * np_, dtype=torch.int32, device="cuda"); g_q = torch.empty(world * qd.numel(), dtype=torch.int32, device="cuda")
            g_t = torch.empty_like(g_q); g_d = torch.empty(world * dd.numel(), dtype=torch.float32, device="cuda")
        run(); barrier()
        lm0 = _lib.kernel_launch_count()
        e0.record(); run(); e1.record(); barrier()
        mms = max_over_ranks(e0.elapsed_time(e1))
        m_launch = _lib.kernel_launch_count() - lm0
        pairs_s = len(pairs_all) / (mms * 1e-3)
        flops = 2.0 * n_feat * n_feat * kdim
        # e2e: host descriptors per pair (H2D 2 x n x k x 4 B, D2H the match list)
        import mavmap_b200 as mm
        t0 = time.perf_counter()
        for (i, j) in mine[: max(2, min(6, np_))]:
            mm.match_brute_force(None, desc[i], None, desc[j], True, 0.9, -1)
        m_e2e = max_over_ranks((time.perf_counter() - t0) / max(2, min(6, np_)))
        secondary = {"metric": "image_pairs_matched_per_sec", "value": pairs_s, "unit": "pairs/s", "ms_per_pair": mms / max(np_, 1),
                     "config": {"workload": "5000 x 5000 SURF-%d descriptors per pair, ratio 0.9 + cross-check, %d pairs/rank" % (kdim, np_)},
                     "impl": "simt (exact fp64-accumulate CUDA cores)" if os.environ.get("MM_MATCH_NO_TC") else "tcgen05 TF32 candidate GEMM + exact re-rank", "gpu_launches": int(m_launch),
                     "algorithmic_tflops": pairs_s * flops / 1e12, "tensor_roofline_frac_of_bf16_peak": pairs_s * flops / 1e12 / peaks["bf16_tflops"],
                     "e2e": {"value": world / m_e2e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * n_feat * kdim * 4, "d2h_bytes_per_step": 12 * 3000}}
        ms_set.close()

    # ------------------------------------------------------------------ the north_star problem (configs[3]: 5000 cams / 10 M obs), N = 1
    big = None
    if rank == 0 and world == 1 and not args.no_cfg4 and args.workload == "ba_cfg2":
        try:
            fbig, _ = synthetic.make_ba_problem(**dict(synthetic.BA_CONFIGS["cfg4"]))
            sb = BASession(fbig.copy(), options(W + K), stream=stream)
            sb.iterate(W); torch.cuda.synchronize()
            e0.record(); nb = sb.iterate(K); e1.record(); torch.cuda.synchronize()
            msb = e0.elapsed_time(e1)
            k1b = sb.time_kernel(0, 20); k2b = sb.time_kernel(1, 5)
            sumb = sb.summary().as_dict(); nblk_b = sb.num_blocks(); cdim_b = sb.coarse_dim(); sb.close()
            k1b_bytes = 184.0 * fbig.n_obs + 48.0 * fbig.n_img + 24.0 * fbig.n_pt
            k2b_bytes = 168.0 * fbig.n_obs + 72.0 * fbig.n_pt + 216.0 * fbig.n_img + 288.0 * nblk_b
            big = {"workload": "cfg4: %d images, %d points, %d observations" % (fbig.n_img, fbig.n_pt, fbig.n_obs), "value": nb / (msb * 1e-3), "unit": UNIT,
                   "ms_per_step": msb / max(nb, 1), "steps": nb, "pcg_iterations": int(sum(sumb["trace_linear_iterations"][W + 1:W + 1 + K])), "coarse_unknowns": cdim_b,
                   "roofline_K1": {"bound": "hbm", "ms": k1b, "achieved": k1b_bytes / (k1b * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": k1b_bytes / (k1b * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                   "roofline_K2": {"bound": "hbm", "ms": k2b, "achieved": k2b_bytes / (k2b * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": k2b_bytes / (k2b * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
            if not args.no_cpu:
                ips_b, dt_b, n_b, cores_b, _ = cpu_ba_iterations(fbig, 2)
                big["cpu_baseline"] = {"value": ips_b, "unit": UNIT, "cores": cores_b, "kind": "port", "sample": "full cfg4 problem, %d LM iterations of the oracle, %.1f s" % (n_b, dt_b)}
            del fbig
        except Exception as e:          # the headline line must not depend on the extra measurement
            big = {"error": repr(e)}

    # ------------------------------------------------------------------ pose_refinement latency, rank 0
    pose_lat = None
    if rank == 0:
        import mavmap_b200 as mmod
        from mavmap_b200.synthetic import _rodrigues, project
        rng = np.random.default_rng(5); npts = 2000
        Xp = rng.uniform([-3, -3, 5], [3, 3, 12], (npts, 3)); rv, tv = np.array([0.08, -0.05, 0.03]), np.array([0.2, -0.3, 0.4])
        prm = list(synthetic.INTRINSICS[1]) + [1]
        uvp = project(1, np.array(prm[:-1]), Xp @ _rodrigues(rv)[0].T + tv) + rng.normal(0, 0.4, (npts, 2))
        po = mmod.BundleAdjustmentOptions(print_summary=False, max_num_iterations=10, function_tolerance=0, gradient_tolerance=0)
        msk = np.ones(npts, bool)
        mmod.pose_refinement(rv + 0.03, tv - 0.08, prm, uvp, Xp, msk, po)
        t0 = time.perf_counter()
        for _ in range(20):
            mmod.pose_refinement(rv + 0.03, tv - 0.08, prm, uvp, Xp, msk, po)
        pose_lat = {"metric": "pose_refinement_latency", "value": (time.perf_counter() - t0) / 20 * 1e6, "unit": "us per call (host buffers in, pose out)",
                    "config": {"workload": "%d 2D-3D pairs, PINHOLE, 10 LM iterations, single-CTA kernel" % npts}}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ips, dt, n, cores, s = cpu_ba_iterations(flat, args.cpu_steps)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "full %s problem, %d LM iterations of the oracle (restated Ceres-semantics LM + Schur + skyline Cholesky, OpenMP; NOT Ceres), %.1f s" % (name, n, dt)}
        # parity of the two engines on this very workload after the same iteration count
        g = flat.copy(); sg = solve_flat(g, options(args.cpu_steps))
        cpu["parity_rel_cost_diff_after_%d_iters" % n] = abs(sg.final_cost - s.final_cost) / s.final_cost
        if secondary is not None:
            secondary["cpu_baseline"] = {"value": cpu_match_pairs(desc, 3), "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
                                         "sample": "3 pairs, cv2.BFMatcher knnMatch x2 + ratio + cross-check (OpenCV %s)" % __import__("cv2").__version__}

    if rank == 0:
        dominant = max((roof_k1, roof_k2), key=lambda r: r["ms"])
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %d images, %d points, %d observations, PINHOLE fx=fy=1000, fixed intrinsics, Cauchy loss, image0 FIXED / image1 FIXED_X" % (name, n_img, n_pt, n_obs),
                           "parallelism": ("one problem, points sharded across %d GPUs, all-reduce of the reduced system per Schur assembly" % world) if sharded else
                                          ("replicas only (one independent BA per GPU)" if world > 1 else "single GPU"),
                           "l2_policy": "inputs larger than L2: %.0f MB of Jacobian records + %.0f MB of observations per LM iteration" % (160.0 * n_obs / 1e6, 24.0 * n_obs / 1e6),
                           "pcg_tolerance": 1e-13, "reduced_system_blocks": n_blocks,
                           "pcg_preconditioner": ("two-level: block-Jacobi + %d similarity-mode coarse unknowns" % coarse_dim) if coarse_dim else "block-Jacobi"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": dict(roof_k1, note="K1 is the kernel north_star names; share of the step is in `breakdown`"),
                "roofline_other": [roof_k2, roof_k3, roof_k3p],
                "breakdown": {"ms_linearize_K1": lin_ms["linearize"], "ms_schur_K2": lin_ms["schur"], "ms_pcg_K3": lin_ms["pcg"], "ms_update_K4": lin_ms["update"],
                              "pcg_iterations_in_timed_steps": int(pcg_iters), "k1_ms": k1_ms, "k2_ms": k2_ms, "k4_cost_ms": k4_ms, "pcg_spmv_ms": spmv_ms,
                              "coarse_setup_ms": coarse_ms, "ms_pcg_per_iteration": lin_ms["pcg"] / max(int(pcg_iters), 1),
                              "dominant_by_time": "K3 PCG" if lin_ms["pcg"] > max(lin_ms["schur"], lin_ms["linearize"]) else dominant["kernel"]},
                "cpu_baseline": cpu, "secondary": secondary, "tertiary": pose_lat, "north_star_cfg4": big, "final_cost": summ["final_cost"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
