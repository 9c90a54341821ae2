"""Turn gpurun_out/*.ncu-rep and the ncu launch list into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py raw   gpurun_out/prof_flow.ncu-rep  > profiles/r01_flow_raw.txt
    python tools/summarize_ncu.py stall gpurun_out/prof_flow.ncu-rep  > profiles/r01_flow_stalls.txt
    python tools/summarize_ncu.py launches gpurun_out/launches_bench.csv > profiles/r01_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__inst_executed.avg.per_cycle_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def raw(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k} = {r[i]} {units[i]}")


def stall(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks[:1]:
        hdr, data = b["rows"][0], b["rows"][1:]
        ix = {h: i for i, h in enumerate(hdr)}
        ns, ie = ix["# Samples"], ix["Instructions Executed"]
        total_s = sum(int(r[ns]) for r in data) or 1
        total_i = sum(int(r[ie]) for r in data)
        print("kernel:", b["name"])
        print(f"warp instructions executed: {total_i}; stall samples: {total_s}")
        for c in [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]:
            s = sum(int(r[ix[c]]) for r in data)
            if s >= 0.01 * total_s:
                print(f"  {c:28s} {100.0 * s / total_s:5.1f} %")
        mix = collections.Counter()
        for r in data:
            t = r[1].strip().split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            mix[op.split(".")[0]] += int(r[ie])
        print("dynamic instruction mix (warp instructions):")
        for op, n in mix.most_common(14):
            print(f"  {op:10s} {n:12d} {100.0 * n / max(total_i, 1):5.1f} %")
        print("hottest SASS lines by stall samples:")
        for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][ns]))[:12]):
            print(f"  [{k:4d}] samples {data[k][ns]:>6s} exec {data[k][ie]:>9s}  {data[k][1].strip()[:80]}")


def launches(path):
    agg = collections.OrderedDict()
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("ffmp::<unnamed>::", "")
        key = (name, r["Grid Size"], r["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"{'kernel':70s} {'grid':>14s} {'block':>12s} {'launches':>8s} {'avg us':>9s} {'share':>7s}")
    for (name, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {grid:>14s} {block:>12s} {n:8d} {ns / n / 1e3:9.2f} {100.0 * ns / total:6.1f}%")


if __name__ == "__main__":
    {"raw": raw, "stall": stall, "launches": launches}[sys.argv[1]](sys.argv[2])
