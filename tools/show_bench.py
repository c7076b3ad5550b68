"""Print the interesting parts of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
for k in ["value", "ms_per_step", "n_gpus", "scaling", "e2e", "parity", "cpu_baseline", "solver", "breakdown", "gpu_launches", "clocks", "replicas"]:
    print(k, d.get(k))
print("roofline", {k: d["roofline"].get(k) for k in ["ms", "frac", "share_of_step", "traffic"]})
for r in d["roofline_other"]:
    print("  ", r["kernel"], "ms", r["ms"], "frac", r["frac"], "share", r.get("share_of_step"))
c = d.get("secondary_cfg2")
if c: print("cfg2", {k: c[k] for k in c if k != "workload"})
c = d.get("secondary_cfg5")
if c: print("cfg5", {k: c[k] for k in c if k != "workload"})
if d.get("secondary"): print("matching", d["secondary"]["value"], d["secondary"]["e2e"], d["secondary"].get("cpu_baseline"))
print("pose", d.get("tertiary"))
