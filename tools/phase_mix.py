"""Split the warp instructions and stall samples of one ncu capture by how often each SASS line ran (a proxy for the phase
it belongs to: per level, per row, per widen iteration ...).    python tools/phase_mix.py rep.ncu-rep [grids]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
grids = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
buckets = collections.OrderedDict()
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name":
        break
    e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    per = e / grids
    key = ("<=1.5/grid" if per <= 1.5 else "<=5/grid (per row)" if per <= 5 else "<=40/grid (per 32-step loop)" if per <= 40
           else "<=64/grid" if per <= 64 else "per level (>64/grid)")
    b = buckets.setdefault(key, [0, 0, 0])
    b[0] += e; b[1] += s; b[2] += 1
    tot_i += e; tot_s += s
for k, (e, s, n) in buckets.items():
    print(f"{k:32s} sass lines {n:5d}  warp-instr {e:11d} ({100.0*e/tot_i:5.1f} %)  per grid {e/grids:8.0f}  samples {100.0*s/max(1,tot_s):5.1f} %")
print("total", tot_i, "per grid", tot_i / grids)
