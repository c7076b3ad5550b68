"""Turn the raw outputs of tools/gpu_round.sh (gpurun_out/) into the committed evidence under profiles/.
usage: python tools/make_profiles.py r1h"""
import csv, json, os, sys, collections
tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
traffic = {}
out = ["# ncu --set full --clock-control none, one launch of every kernel (tools/gpu_round.sh %s); per-launch values" % tag, ""]
for name in ("prof_ba_cfg2", "prof_ba_cfg4", "prof_match"):
    fn = os.path.join(G, "%s_%s.csv" % (name, tag))
    if not os.path.exists(fn): continue
    rows = list(csv.reader(open(fn))); hdr, units = rows[0], rows[1]; kn = hdr.index("Kernel Name")
    out.append("## %s" % name)
    for r in rows[2:]:
        k = r[kn].split("(")[0].replace("void ", "").replace("mm::", "")
        vals = []
        for metric, short in WANT:
            if metric in hdr:
                i = hdr.index(metric); vals.append("%s=%s%s" % (short, r[i], (" " + units[i]) if units[i] and units[i] != "%" else ("%" if units[i] == "%" else "")))
        out.append("%-36s %s" % (k[:36], "  ".join(vals)))
        if name.startswith("prof_ba"):
            cfg = name.split("_")[2]
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]); wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            traffic.setdefault(cfg, {})[k.replace(" ", "")] = rd + wr
    out.append("")
if traffic:            # only when this round has `--set full` captures
    open(os.path.join(P, "ncu_full_%s.txt" % tag), "w").write("\n".join(out))
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
# launch list: aggregate by kernel
fn = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(fn):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
    hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[mv].replace(",", ""))
        except ValueError: continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
        k = r[kn][:72]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 python bench.py --steps 2 --warmup 3 --no-cpu   (%s)" % tag,
             "# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes", ""]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-72s n=%5d total %10.1f us  avg %9.2f us  %5.1f%%" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    open(os.path.join(P, "launches_%s.txt" % tag), "w").write("\n".join(lines) + "\n")
for f in ("", "_cfg2", "_cfg4", "_ref"):
    src = os.path.join(G, "bench_%s%s.json" % (tag, f))
    if os.path.exists(src): open(os.path.join(P, "bench_%s%s.json" % (tag, f)), "w").write(open(src).read())
print(open(os.path.join(P, "ncu_traffic.json")).read()[:600])
