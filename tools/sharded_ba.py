"""Multi-GPU (point-sharded) bundle adjustment against the single-GPU engine on the same problem.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_ba.py mid cfg2
Every rank checks cost trace / step pattern / parameters against an unsharded solve on its own GPU and rank 0 prints
the timings.  Exit code 1 on any mismatch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mavmap_b200 import synthetic
from mavmap_b200.ba import BASession, default_c_options
from mavmap_b200.parallel import make_allreduce_callback

CFG = dict(synthetic.BA_CONFIGS)
CFG["mid"] = dict(n_img=120, n_obs_target=120000, track_len=4, seed=777)
CFG["midrefine"] = dict(n_img=120, n_obs_target=120000, track_len=4, seed=779, models=[1, 2], refine_camera_params=True)      # refined intrinsics of a two-camera rig, sharded


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cb = make_allreduce_callback()
    ok = True
    for name in sys.argv[1:] or ["mid"]:
        iters = 8
        flat, _ = synthetic.make_ba_problem(**CFG[name])
        o = default_c_options(); o.max_num_iterations = iters; o.function_tolerance = 0; o.gradient_tolerance = 0
        stream = torch.cuda.current_stream().cuda_stream
        def run(sharded):
            f = flat.copy()
            s = BASession(f, o, stream=stream, rank=rank if sharded else 0, world=world if sharded else 1, allreduce=cb if sharded else None)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); n = s.iterate(iters); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            d = s.summary().as_dict(); s.download(); s.close()
            return f, d, ms, n
        fs, ds, ms_s, n_s = run(True)
        f1, d1, ms_1, n_1 = run(False)
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
        cost_rel = max(abs(a - b) / b for a, b in zip(ds["trace_cost"], d1["trace_cost"]))
        good = ds["trace_accepted"] == d1["trace_accepted"] and cost_rel < 1e-9 and rel(fs.poses, f1.poses) < 1e-6 and rel(fs.pts, f1.pts) < 1e-5 and rel(fs.intr, f1.intr) < 1e-6
        t = torch.tensor([ms_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and good
        print("[rank %d] %s: %d imgs %d obs | sharded x%d: %d it in %.1f ms (%.1f it/s, pcg %s) | single: %.1f ms (%.1f it/s) | cost rel %.1e poses %.1e pts %.1e accepted-equal %s -> %s" % (
            rank, name, flat.n_img, flat.n_obs, world, n_s, float(t.item()), 1e3 * n_s / float(t.item()), ds["trace_linear_iterations"][1:4], ms_1, 1e3 * n_1 / ms_1,
            cost_rel, rel(fs.poses, f1.poses), rel(fs.pts, f1.pts), ds["trace_accepted"] == d1["trace_accepted"], "OK" if good else "MISMATCH"), flush=True)
        if rank == 0:
            print("   breakdown sharded ms:", {k: round(v, 2) for k, v in ds["ms"].items()}, "single:", {k: round(v, 2) for k, v in d1["ms"].items()}, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
