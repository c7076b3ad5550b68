#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MAVMAP hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: the oracle restatement of the
                                                            reference's Ceres path, all host cores)

Primary metric (BASELINE.json): bundle-adjustment LM iterations per second on the configuration the
metric is quoted on, configs[3] = "cfg4": 5000 cameras / 2 M points / 10 M observations.  A "step" is one LM
iteration = residual+Jacobian (K1) + Schur assembly (K2) + PCG solve preconditioned by the sparse tile Cholesky
(K3) + back-substitution / candidate cost (K4) + accept/reject.  cfg2 (500 images) and the matching
throughput (image pairs/s, 5k x 5k SURF-64) ride along as `secondary_cfg2` / `secondary`.

value    : inputs resident in HBM (mm_ba_session_*), CUDA-event timed on the session's stream
e2e      : same metric through mm_ba_solve with HOST buffers (upload + structure setup + K
           iterations + download inside the timed region)
parity   : computed in the run: cost, step pattern and parameters against the CPU oracle after the same
           number of LM iterations on this very problem
N > 1    : ONE cfg4 problem with its points sharded across the N GPUs (strong scaling; the reduced system is
           all-reduced once per Schur assembly and solved on every rank); `replicas` (N independent
           problems, what north_star prescribes for BA) is reported beside it.
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
os.environ.setdefault("MM_BA_DIAG", "1")          # the engine reports every refused LM step on stderr (none is expected on these workloads)

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


WORKLOADS = {"ba_cfg2": "cfg2", "ba_cfg4": "cfg4", "ba_cfg1": "cfg1", "ba_small": "small"}


def make_problem(name, seed_offset=0):
    from mavmap_b200 import synthetic
    kw = dict(synthetic.BA_CONFIGS[name])
    kw["seed"] = kw["seed"] + seed_offset
    flat, _ = synthetic.make_ba_problem(**kw)
    return flat


def workload_string(name, flat):
    """identical in both arms (the driver compares it)"""
    return ("%s: %d images, %d points, %d observations, PINHOLE fx=fy=1000, fixed intrinsics, Cauchy loss, image 0 FIXED / image 1 FIXED_X"
            % (name, flat.n_img, flat.n_pt, flat.n_obs))


def options(iters, oracle=False):
    from mavmap_b200.ba import default_c_options
    if oracle:
        from oracle import orc
        o = orc.default_options()
    else:
        o = default_c_options()
    o.max_num_iterations = iters; o.function_tolerance = 0.0; o.gradient_tolerance = 0.0     # fixed iteration count
    return o


def cpu_ba(flat, iters):
    """Oracle (restated Ceres-semantics LM + Schur + Cholesky, OpenMP) on the host cores: returns (seconds, summary, solved copy)."""
    from oracle import orc
    f = flat.copy()
    t = time.perf_counter()
    s = orc.solve_flat(f, options(iters, oracle=True))
    return time.perf_counter() - t, s, f


def parity_block(g, sg, c, sc, iters):
    """GPU result (problem g, summary dict sg) against the oracle's (c, sc) after the same iteration count"""
    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0
    n = min(len(sg["trace_cost"]), len(sc["trace_cost"]))
    tc_g, tc_c = np.array(sg["trace_cost"][:n]), np.array(sc["trace_cost"][:n])
    return {"against": "CPU oracle (oracle/orc_ba.c: restated Ceres LM + Schur + direct Cholesky), same problem, %d LM iterations" % iters,
            "rel_final_cost_diff": float(abs(sg["final_cost"] - sc["final_cost"]) / sc["final_cost"]),
            "max_rel_cost_trace_diff": float(np.max(np.abs(tc_g - tc_c) / tc_c)),
            "same_step_pattern": sg["trace_accepted"][:n] == sc["trace_accepted"][:n],
            "max_rel_pose_diff": rel(g.poses, c.poses), "max_rel_point_diff": rel(g.pts, c.pts),
            "tolerance_north_star": 1e-6,
            "max_pcg_iterations_per_solve": int(max(sg["trace_linear_iterations"]) if sg["trace_linear_iterations"] else 0)}


def cpu_match_pairs(desc, n_pairs):
    import cv2
    from oracle import orc
    cv2.setNumThreads(os.cpu_count() or 1)
    t = time.perf_counter()
    for p in range(n_pairs):
        orc.match_pair_cv2(desc[p % len(desc)], desc[(p + 1) % len(desc)], ratio_test=True, max_ratio=0.9)
    return n_pairs / (time.perf_counter() - t)


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  Real Ceres/OpenCV-C++ cannot be built here (not in the image), so this is the
    oracle port of the reference's BA path on all host cores, on the same problem, for W + K LM iterations; the rate is
    taken over the last K."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1: undo it before libgomp loads
    from oracle import orc
    name = WORKLOADS[args.workload]
    flat = make_problem(name)
    W, K = args.warmup, args.steps
    # two runs of the same deterministic trajectory: W iterations (untimed part), then W + K; the difference is K steps
    t_w = 0.0
    if W > 0:
        t_w, _, _ = cpu_ba(flat, W)
    t_all, s, _ = cpu_ba(flat, W + K)
    n = s.num_successful_steps + s.num_unsuccessful_steps
    dt = max(t_all - t_w, 1e-9)
    ips = K / dt
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / max(K, 1), "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_string(name, flat)},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                             "sample": "full %s problem: %d LM iterations of the oracle timed as (run of %d) - (run of %d) (restated Ceres-semantics LM + Schur + skyline Cholesky, OpenMP; NOT Ceres: no Ceres in the image)" % (name, K, n, W)},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "final_cost": s.final_cost}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ba_cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--match-pairs", type=int, default=120)
    ap.add_argument("--match-images", type=int, default=200, help="images of the resident descriptor set of the matching block")
    ap.add_argument("--cpu-steps", type=int, default=3, help="LM iterations of the CPU oracle for cpu_baseline and the parity block")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--no-cfg2", action="store_true", help="skip the secondary 500-image and 1000-image-rig measurements")
    ap.add_argument("--ba-mode", default="auto", choices=["auto", "replicas", "sharded"],
                    help="N > 1: ONE problem with its points sharded (auto / sharded: strong scaling) or independent problems per GPU")
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

    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    W, K = args.warmup, args.steps
    name = WORKLOADS[args.workload]
    sharded = world > 1 and args.ba_mode != "replicas"

    def timed_session(flat, make_session):
        """W warm-up + K timed LM iterations on a resident session; returns (ms over K, launches, summary dict, session)"""
        sess = make_session(flat.copy())
        sess.iterate(W)
        barrier()
        ms_w = sess.summary().as_dict()["ms"]                 # phase clocks after the warm-up: the timed region's share is the difference
        l0 = _lib.kernel_launch_count()
        e0.record()
        done = sess.iterate(K)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        assert done == K, "LM stopped early (%d of %d iterations)" % (done, K)
        summ = sess.summary().as_dict()
        summ["ms_timed"] = {k: summ["ms"][k] - ms_w[k] for k in ("linearize", "schur", "pcg", "update")}
        return ms, _lib.kernel_launch_count() - l0, summ, sess

    # ------------------------------------------------------------------ headline BA problem, resident
    flat = make_problem(name)                                       # the same problem on every rank
    rep = flat if rank == 0 else (make_problem(name, 1000 * rank) if world > 1 else flat)     # one independent problem per rank (replicas, host-buffer call)
    if sharded:
        from mavmap_b200.parallel import make_allreduce_callback
        ar_cb = make_allreduce_callback()
        mk = lambda f: BASession(f, options(W + K), stream=stream, rank=rank, world=world, allreduce=ar_cb)
    else:
        mk = lambda f: BASession(f, options(W + K), stream=stream)
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.25)                                       # nvidia-smi needs a moment to produce its first sample
    t_load0 = time.time()
    ms, launches, summ, sess = timed_session(flat if sharded else rep, mk)
    clocks = sampler.stop(t_load0, time.time())
    value = (1 if sharded else world) * K / (ms * 1e-3)
    n_blocks = sess.num_blocks(); info = sess.solver_info()

    # per-kernel timings on the resident state (live, CUDA events inside the library)
    n_obs, n_pt, n_img = flat.n_obs, flat.n_pt, flat.n_img
    share = 1.0 / world if sharded else 1.0                                          # observations / points per rank
    k1_ms = sess.time_kernel(0, 20); k2_ms = sess.time_kernel(1, 10); k4_ms = sess.time_kernel(2, 20)
    k1_bytes = (184.0 * n_obs + 24.0 * n_pt) * share + 48.0 * n_img              # K1 algorithmic bytes
    k2_bytes = (168.0 * n_obs + 72.0 * n_pt) * share + 216.0 * n_img + 288.0 * n_blocks
    traffic = None                       # dram bytes per launch of K1 from the committed `ncu --set full` capture of this workload (not measured in this run)
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp) and world == 1:
        tj = json.load(open(tp)).get(name, {})
        traffic = next((v for k, v in sorted(tj.items()) if k.startswith("k_residual_jacobian<1,0")), None)     # the Jacobian pass (any cost variant)
    lin_ms = summ["ms_timed"]
    pcg_iters = summ["trace_linear_iterations"][W + 1:W + 1 + K]
    roof_k1 = {"kernel": "k_residual_jacobian (K1)", "bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full of this workload, committed; not re-measured in this run)" if traffic else None,
               "ms": k1_ms, "peak_source": peaks["source"], "algorithmic_bytes": k1_bytes,
               "share_of_step": k1_ms / (ms / K)}
    roof_k1["frac"] = roof_k1["achieved"] / roof_k1["peak"]
    roof_k2 = {"kernel": "Schur assembly (K2: k_schur_point + k_schur_blocks + k_schur_cam)", "bound": "hbm", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "traffic": None, "ms": k2_ms, "share_of_step": k2_ms / (ms / K)}
    roof_k2["frac"] = roof_k2["achieved"] / roof_k2["peak"]
    others = [roof_k2]
    solver = {"preconditioner": info["preconditioner"], "pcg_iterations_per_solve": {"min": int(min(pcg_iters)), "max": int(max(pcg_iters)), "mean": float(np.mean(pcg_iters))}}
    if info["preconditioner"] == "sparse tile Cholesky":
        f_ms = sess.time_kernel(5, 5); a_ms = sess.time_kernel(6, 5)
        sm_mhz = clocks.get("sm_max_mhz") or 1965.0
        fp64_peak = 148 * 128 * sm_mhz * 1e6 / 1e12                                  # 64 FMA / clk / SM, nominal: no measured fp64 peak in MEASURED_PEAKS.json
        solver.update({"tiles_of_L": info["tiles"], "tile_mb": info["tile_mb"], "tile_products": info["tile_products"], "gflop_per_factorisation": info["flops"] / 1e9,
                       "ms_assembly_and_factorisation": f_ms, "ms_per_application": a_ms, "substitution_tasks": info["substitution_tasks"]})
        others.append({"kernel": "k_tc_factor (K3: sparse tile Cholesky; dominant by time)", "bound": "fp64", "achieved": info["flops"] / (f_ms * 1e-3) / 1e12, "peak": fp64_peak,
                       "unit": "TFLOP/s", "traffic": None, "ms": f_ms, "frac": info["flops"] / (f_ms * 1e-3) / 1e12 / fp64_peak, "share_of_step": f_ms / (ms / K),
                       "note": "dependency-bound (critical path through the dense separator blocks), not pipe-bound; peak = 148 SMs x 64 fp64 FMA/clk x max SM clock (nominal)"})
    sess.close()

    # ------------------------------------------------------------------ same problem, end to end through mm_ba_solve (host buffers)
    fe = rep                                                  # the host-buffer call is per GPU: one problem each
    warm = fe.copy(); solve_flat(warm, options(1))            # load kernels / allocator warm-up, untimed
    barrier()
    f2 = fe.copy()
    t0 = time.perf_counter()
    s2 = solve_flat(f2, options(K))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    n2 = s2.num_successful_steps + s2.num_unsuccessful_steps
    h2d = 24 * fe.n_obs + 48 * fe.n_img + 24 * fe.n_pt + 72 + 6 * fe.n_img + fe.n_pt
    d2h = 48 * fe.n_img + 24 * fe.n_pt + 72
    e2e = {"value": world * n2 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(n2, 1)), "d2h_bytes_per_step": int(d2h / max(n2, 1)),
           "total_ms": 1e3 * e2e_s, "setup_ms": s2.ms_setup, "steps": n2,
           "note": "mm_ba_solve on host arrays%s: upload + device structure setup + host symbolic analysis of the tile Cholesky + %d LM iterations + download" % (" (one independent problem per GPU)" if world > 1 else "", n2)}

    # ------------------------------------------------------------------ parity + CPU baseline on this very problem (rank 0, N = 1)
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import orc
        dt, sc, c = cpu_ba(flat, args.cpu_steps)
        n = sc.num_successful_steps + sc.num_unsuccessful_steps
        cpu = {"value": n / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
               "sample": "full %s problem, %d LM iterations of the oracle (restated Ceres-semantics LM + Schur + skyline Cholesky, OpenMP; NOT Ceres), %.1f s" % (name, n, dt)}
        g = flat.copy(); sg = solve_flat(g, options(args.cpu_steps)).as_dict()
        parity = parity_block(g, sg, c, sc.as_dict(), args.cpu_steps)

    # ------------------------------------------------------------------ replicas beside the sharded line (N > 1)
    replicas = None
    if sharded:
        msr, _, _, sr = timed_session(rep, lambda f: BASession(f, options(W + K), stream=stream))
        sr.close()
        replicas = {"value": world * K / (msr * 1e-3), "unit": UNIT, "scaling": "weak", "ms_per_step": msr / K,
                    "note": "N independent %s problems, one per GPU, no communication (what north_star prescribes for BA that fits one HBM)" % name}

    # ------------------------------------------------------------------ secondary BA configuration: cfg2 (N = 1)
    cfg2 = None
    if rank == 0 and world == 1 and not args.no_cfg2 and name != "cfg2":
        try:
            f2c = make_problem("cfg2")
            ms2, _, sum2, sb = timed_session(f2c, lambda f: BASession(f, options(W + K), stream=stream))
            k1b = sb.time_kernel(0, 20); k2b = sb.time_kernel(1, 10); nblk2 = sb.num_blocks(); sb.close()
            wb = f2c.copy(); solve_flat(wb, options(1))
            fb = f2c.copy(); t0 = time.perf_counter(); s2b = solve_flat(fb, options(K)); torch.cuda.synchronize(); e2b = time.perf_counter() - t0
            cfg2 = {"workload": workload_string("cfg2", f2c), "value": K / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K, "steps": K,
                    "e2e": {"value": (s2b.num_successful_steps + s2b.num_unsuccessful_steps) / e2b, "unit": UNIT, "total_ms": 1e3 * e2b, "setup_ms": s2b.ms_setup},
                    "pcg_iterations": sum2["trace_linear_iterations"][W + 1:W + 1 + K], "breakdown_ms": sum2["ms_timed"],
                    "roofline_K1": {"bound": "hbm", "ms": k1b, "frac": (184.0 * f2c.n_obs + 48.0 * f2c.n_img + 24.0 * f2c.n_pt) / (k1b * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                    "roofline_K2": {"bound": "hbm", "ms": k2b, "frac": (168.0 * f2c.n_obs + 72.0 * f2c.n_pt + 216.0 * f2c.n_img + 288.0 * nblk2) / (k2b * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
            if not args.no_cpu:
                dtc, scc, cc = cpu_ba(f2c, args.cpu_steps)
                cfg2["cpu_baseline"] = {"value": (scc.num_successful_steps + scc.num_unsuccessful_steps) / dtc, "unit": UNIT, "kind": "port", "sample": "full cfg2 problem, %d LM iterations of the oracle, %.1f s" % (args.cpu_steps, dtc)}
                gg = f2c.copy(); sgg = solve_flat(gg, options(args.cpu_steps)).as_dict()
                cfg2["parity"] = parity_block(gg, sgg, cc, scc.as_dict(), args.cpu_steps)
        except Exception as e:          # the headline line must not depend on the extra measurement
            cfg2 = {"error": repr(e)}

    # ------------------------------------------------------------------ cfg5-sized rig with refined intrinsics (N = 1)
    cfg5 = None
    if rank == 0 and world == 1 and not args.no_cfg2 and name != "cfg5":
        try:
            f5 = make_problem("cfg5")
            ms5, _, sum5, s5 = timed_session(f5, lambda f: BASession(f, options(W + K), stream=stream))
            s5.close()
            cfg5 = {"workload": "cfg5-sized: %d images of a PINHOLE + OPENCV rig, %d points, %d observations, refine_camera_params (2 x 9 intrinsics in the border of the reduced system), global BA only (no mapper loop)" % (f5.n_img, f5.n_pt, f5.n_obs),
                    "value": K / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5 / K, "steps": K,
                    "pcg_iterations": sum5["trace_linear_iterations"][W + 1:W + 1 + K], "breakdown_ms": sum5["ms_timed"]}
            if not args.no_cpu:
                dtc, scc, cc = cpu_ba(f5, args.cpu_steps)
                cfg5["cpu_baseline"] = {"value": (scc.num_successful_steps + scc.num_unsuccessful_steps) / dtc, "unit": UNIT, "kind": "port", "sample": "full problem, %d LM iterations of the oracle, %.1f s" % (args.cpu_steps, dtc)}
                gg = f5.copy(); sgg = solve_flat(gg, options(args.cpu_steps)).as_dict()
                cfg5["parity"] = parity_block(gg, sgg, cc, scc.as_dict(), args.cpu_steps)
                cfg5["parity"]["max_rel_intrinsics_diff"] = float(np.max(np.abs(gg.intr - cc.intr) / np.maximum(np.abs(cc.intr), 1.0)))
        except Exception as e:
            cfg5 = {"error": repr(e)}

    # ------------------------------------------------------------------ matching (secondary), sharded by pair
    secondary = None
    if not args.no_match:
        n_feat, kdim, n_imgs = 5000, 64, args.match_images
        desc, xy = synthetic.make_descriptors(n_imgs, n_feat, kdim, seed=0xF00D + 3)
        ms_set = MatchSet(desc, None)                      # the whole sequence resident: descriptors + both TF32 operand copies
        # the mapper's pair pattern: every image against its two predecessors (sequential_mapper.cc process()), cycled
        pairs_seq = [(i, i + 1) for i in range(n_imgs - 1)] + [(i, i + 2) for i in range(n_imgs - 2)]
        pairs_all = (pairs_seq * (1 + args.match_pairs * world // len(pairs_seq)))[: args.match_pairs * world]
        mine = pairs_all[rank::world]
        np_ = len(mine)
        cnt = torch.zeros(np_, dtype=torch.int32, device="cuda"); qd = torch.empty(np_ * n_feat, dtype=torch.int32, device="cuda")
        td = torch.empty_like(qd); dd = torch.empty(np_ * n_feat, dtype=torch.float32, device="cuda")
        col = torch.arange(n_feat, device="cuda", dtype=torch.int32)[None, :]
        gathered = {}
        def run():
            ms_set.match_pairs_device(mine, cnt.data_ptr(), qd.data_ptr(), td.data_ptr(), dd.data_ptr(), n_feat, stream, True, 0.9, -1)
            if world > 1:
                # the one exchange step of the sharded path, ragged: counts first, then the PACKED match lists (about a third of
                # the padded slots) - uneven all_gather over NCCL/NVLink
                dist.all_gather_into_tensor(g_cnt, cnt)
                keep = (col < cnt[:, None]).reshape(-1)
                packed = torch.stack([qd[keep], td[keep], dd[keep].view(torch.int32)])           # [3, matches of this rank]
                sizes = g_cnt.view(world, np_).sum(dim=1).tolist()                              # (host read of N numbers: the sizes of the ragged gather)
                outs = [torch.empty((3, int(n)), dtype=torch.int32, device="cuda") for n in sizes]
                dist.all_gather(outs, packed)
                gathered["matches"] = sum(sizes); gathered["bytes"] = 12 * sum(sizes)
        if world > 1:
            g_cnt = torch.empty(world * np_, dtype=torch.int32, device="cuda")
        run(); barrier()
        lm0 = _lib.kernel_launch_count()
        e0.record(); run(); e1.record(); barrier()
        mms = max_over_ranks(e0.elapsed_time(e1))
        m_launch = _lib.kernel_launch_count() - lm0
        pairs_s = len(pairs_all) / (mms * 1e-3)
        flops = 2.0 * n_feat * n_feat * kdim
        # e2e: the one-pair entry point on HOST descriptor arrays, in the mapper's order (image i against i-1 and i-2,
        # sequential_mapper.cc process()); the library keeps its last four images resident, so an image is uploaded when it
        # first appears - the bytes below are counted by the library (mm_match_pair_counters), the D2H is count | q | t | dist
        import mavmap_b200 as mm
        seq = [(i - d, i) for i in range(2, n_imgs) for d in (1, 2)]
        n_e2e = max(2, min(24, np_)); warm_e2e = 4
        host_desc = [np.ascontiguousarray(desc[i]) for i in range(n_imgs)]
        for (i, j) in seq[:warm_e2e]:
            mm.match_brute_force(None, host_desc[i], None, host_desc[j], True, 0.9, -1)
        c0 = _lib.match_pair_counters()
        t0 = time.perf_counter()
        for (i, j) in seq[warm_e2e:warm_e2e + n_e2e]:
            mm.match_brute_force(None, host_desc[i], None, host_desc[j], True, 0.9, -1)
        m_e2e = max_over_ranks((time.perf_counter() - t0) / n_e2e)
        c1 = _lib.match_pair_counters()
        secondary = {"metric": "image_pairs_matched_per_sec", "value": pairs_s, "unit": "pairs/s", "ms_per_pair": mms / max(np_, 1),
                     "config": {"workload": "5000 x 5000 SURF-%d descriptors per pair, ratio 0.9 + cross-check, %d pairs/rank out of a resident %d-image sequence (%.0f MB of descriptors + TF32 operand copies per GPU), image i against i+1 and i+2" % (kdim, np_, n_imgs, n_imgs * n_feat * (kdim + 2 * 96) * 4 / 1e6),
                                "exchange": ("ragged all_gather of counts + packed lists, %d matches = %.1f MB per step" % (gathered.get("matches", 0), gathered.get("bytes", 0) / 1e6)) if world > 1 else "none (single GPU)"},
                     "impl": "simt (exact fp64-accumulate CUDA cores)" if os.environ.get("MM_MATCH_NO_TC") else "tcgen05 TF32 candidate GEMM + exact re-rank", "gpu_launches": int(m_launch),
                     "algorithmic_tflops": pairs_s * flops / 1e12, "algorithmic_tflops_per_gpu": pairs_s * flops / 1e12 / world,
                     "tensor_roofline_frac_of_bf16_peak_per_gpu": pairs_s * flops / 1e12 / world / peaks["bf16_tflops"],
                     "e2e": {"value": world / m_e2e, "unit": "pairs/s", "h2d_bytes_per_step": int((c1[2] - c0[2]) / n_e2e), "d2h_bytes_per_step": 16 + 12 * n_feat,
                             "pairs": n_e2e, "descriptor_arrays_uploaded": int(c1[1] - c0[1]),
                             "note": "mm_match_pair on host arrays, image i against i-1 and i-2; the last four images stay resident in the library"}}
        if rank == 0 and world == 1 and not args.no_cpu:
            secondary["cpu_baseline"] = {"value": cpu_match_pairs(desc, 3), "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
                                         "sample": "3 pairs, cv2.BFMatcher knnMatch x2 + ratio + cross-check (OpenCV %s)" % __import__("cv2").__version__}
        ms_set.close()

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
        try:        # the same problem 64 times through mm_pose_refine_batch: one launch, one CTA per problem
            nb = 64
            args_b = lambda: (np.tile(rv + 0.03, (nb, 1)), np.tile(tv - 0.08, (nb, 1)), [prm] * nb, [uvp] * nb, [Xp] * nb, None, po)
            mmod.pose_refinement_batch(*args_b())
            t0 = time.perf_counter()
            for _ in range(5):
                mmod.pose_refinement_batch(*args_b())
            pose_lat["batch"] = {"problems": nb, "us_per_problem": (time.perf_counter() - t0) / 5 / nb * 1e6,
                                 "note": "mm_pose_refine_batch, host buffers in, poses out (includes packing %d x %d pairs on the host)" % (nb, npts)}
        except Exception as e:
            pose_lat["batch"] = {"error": repr(e)}

    if rank == 0:
        step_ms = ms / K
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(name, flat),
                           "parallelism": ("one problem, points (and with them observations, Jacobian records, K1/K2/K4) sharded across %d GPUs; limiting collective: all-reduce of the reduced "
                                           "system S | rhs | gradient (%.0f MB) once per Schur assembly over NCCL/NVLink; the reduced solve (tile Cholesky + PCG) is replicated" % (world, (36.0 * n_blocks + 18.0 * n_img) * 8 / 1e6)) if sharded else
                                          ("replicas only (one independent BA per GPU)" if world > 1 else "single GPU"),
                           "l2_policy": "inputs larger than L2: %.0f MB of Jacobian records + %.0f MB of observations per LM iteration" % (160.0 * n_obs * share / 1e6, 24.0 * n_obs * share / 1e6),
                           "pcg_tolerance": 1e-13, "reduced_system_blocks": n_blocks, "pcg_preconditioner": info["preconditioner"]},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": dict(roof_k1, note="K1 is the HBM-bound kernel north_star names; the kernel that dominates the step by time is in roofline_other / breakdown"),
                "roofline_other": others,
                "solver": solver,
                "breakdown": {"ms_linearize_K1": lin_ms["linearize"], "ms_schur_K2": lin_ms["schur"], "ms_solve_K3": lin_ms["pcg"], "ms_update_K4": lin_ms["update"],
                              "note": "device ms inside the %d timed LM iterations (CUDA events in the library)" % K,
                              "k1_ms": k1_ms, "k2_ms": k2_ms, "k4_cost_ms": k4_ms,
                              "dominant_by_time": max((("K3 solve", lin_ms["pcg"]), ("K2 Schur assembly", lin_ms["schur"]), ("K1 linearize", lin_ms["linearize"]), ("K4 update", lin_ms["update"])), key=lambda kv: kv[1])[0]},
                "parity": parity, "cpu_baseline": cpu, "secondary": secondary, "secondary_cfg2": cfg2, "secondary_cfg5": cfg5, "replicas": replicas, "tertiary": pose_lat, "final_cost": summ["final_cost"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
