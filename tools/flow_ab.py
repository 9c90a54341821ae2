"""A/B timing of the flow-field operator on 4096 generated 128 x 128 maps (CUDA events, median of 7, 3 warm-ups).
    python tools/flow_ab.py            # runs the variants below, one subprocess each (the switches are read at first use)
    python tools/flow_ab.py child      # one measurement in this process
AB_N / AB_G select the batch and the grid.  (History: profiles/r02a_flow_ab.txt = resident warps per SM, r02d_flow_order_ab.txt =
deepest-first hand-out order against the natural order, which won and stayed.)"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    N = int(os.environ.get("AB_N", "4096"))
    G = int(os.environ.get("AB_G", "128"))
    dev = torch.device("cuda:0")
    gids = torch.arange(N, device=dev)
    occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), G, seed=1234)
    goals = scen[:, 5:7].contiguous()
    out = []
    ws = ffmp.ops.flow_field_workspace(N, G, dev)
    bufs = (torch.empty((N, G, G), dtype=torch.int32, device=dev), torch.empty((N, G, G), dtype=torch.uint8, device=dev))
    goals = goals.to(torch.int32).contiguous()
    REPS = 10           # launches per timing: the Python call (~30 us) overlaps the previous launch
    for want_cost in (True, False):
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(REPS):
                ffmp.ops.flow_field(occ, goals, want_cost=want_cost, out=bufs, workspace=ws)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b) / REPS)
        ts.sort()
        out.append(ts[len(ts) // 2])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ffmp.ops.flow_field(occ, goals); b.record(); torch.cuda.synchronize()
    print(json.dumps({"ms": out[0], "nocost_ms": out[1], "frac_6B": N * G * G * 6 / (out[0] * 1e-3) / 6549.8e9,
                      "single_call_ms": a.elapsed_time(b)}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        variants = []
        for n in ("4096", "4096", "2048", "8192", "16384"):
            variants += [{"AB_N": n}]
        for v in variants:
            env = dict(os.environ, **v)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
            print(v, r.stdout.strip() or r.stderr[-400:], flush=True)
