"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per source line."""
import csv
import sys

path, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = cur = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif r[0].isdigit():
        out.append((cur, int(r[0]), r[1].strip(), dict(zip(hdr, r))))


def num(x):
    try:
        return float(x)
    except Exception:
        return 0.0


tot_s = sum(num(d['Warp Stall Sampling (All Samples)']) for *_, d in out)
tot_i = sum(num(d['Thread Instructions Executed']) for *_, d in out)
print("thread instructions per unit: %.0f" % (tot_i / units))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tt = {h: sum(num(d.get(h, 0)) for *_, d in out) for h in stalls}
T = sum(tt.values())
print("stall mix: " + ", ".join("%s %.0f%%" % (h[6:], 100 * v / T) for h, v in sorted(tt.items(), key=lambda t: -t[1])[:9]))
out.sort(key=lambda t: -num(t[3]['Warp Stall Sampling (All Samples)']))
for f, l, src, d in out[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    top = sorted(((num(d.get(h, 0)), h[6:]) for h in stalls), reverse=True)[:2]
    print("%-13s %4d samp %5.1f%% inst/unit %5.1f  %-64s %s" % (
        f[:13], l, 100 * num(d['Warp Stall Sampling (All Samples)']) / tot_s,
        num(d['Thread Instructions Executed']) / units, src[:64], ",".join(h for _, h in top)))
