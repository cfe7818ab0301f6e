"""Summarises the ncu outputs of tools/profile_bench.sh into profiles/ (text + traffic.json)."""
import csv
import json
import os
import subprocess
import sys

tag, dtype = sys.argv[1], sys.argv[2]  # files gpurun_out/<tag>_<dtype>_{launches.csv,full.ncu-rep,plain.json}
wkey = sys.argv[3] if len(sys.argv) > 3 else "c2_%s" % dtype  # key in profiles/traffic.json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "gpurun_out")
dst = os.path.join(root, "profiles")
os.makedirs(dst, exist_ok=True)
lines = []

# ---- launch list: per-kernel totals
rows = [r for r in csv.reader(open(os.path.join(src, "%s_%s_launches.csv" % (tag, dtype)))) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[rows.index(hdr) + 1:]:
    name = r[ki].split("<")[0].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)  # -> us
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
lines.append("== kernel launch list (ncu --metrics gpu__time_duration.sum, serialised, cold cache): share of GPU time")
for name, (n, us) in sorted(tot.items(), key=lambda t: -t[1][1]):
    lines.append("  %-60s launches %4d  total %12.1f us  share %5.1f%%" % (name[:60], n, us, 100 * us / total))

# ---- full capture of the persistent kernel
rep = os.path.join(src, "%s_%s_full.ncu-rep" % (tag, dtype))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
get = lambda k: (v[h.index(k)], u[h.index(k)]) if k in h else ("n/a", "")  # noqa: E731
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines.append("== ncu --set full, kernel %s" % get("Kernel Name")[0][:100])
for k in keys:
    val, unit = get(k)
    lines.append("  %-66s %s %s" % (k, val, unit))


def to_bytes(val, unit):
    f = float(val.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


traffic = to_bytes(*get("dram__bytes_read.sum")) + to_bytes(*get("dram__bytes_write.sum"))
lines.append("  DRAM traffic per launch (read+write): %.3f GB" % (traffic / 1e9))
tj = os.path.join(dst, "traffic.json")
data = json.load(open(tj)) if os.path.exists(tj) else {}
data[wkey] = traffic
data[wkey + "_source"] = "profiles/%s_%s_ncu_summary.txt" % (tag, dtype)
json.dump(data, open(tj, "w"), indent=1)
bench = open(os.path.join(src, "%s_%s_plain.json" % (tag, dtype))).read().strip()
lines.append("== bench line of the same command (plain run, not under ncu)")
lines.append(bench)
open(os.path.join(dst, "%s_%s_ncu_summary.txt" % (tag, dtype)), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
