"""Large maps (512 x 512, p = 0.3): the one-CTA-per-grid kernel against the thread-block-cluster variant (FFMP_FLOW_CLUSTER=1, two
CTAs per grid) at 512 / 296 / 148 / 64 / 16 grids per launch, and the config-4 step."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    dev = torch.device("cuda:0")
    out = {"cluster": os.environ.get("FFMP_FLOW_CLUSTER", "0")}
    for n in (512, 296, 148, 64, 16):
        gids = torch.arange(n, device=dev)
        occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), 512, p_occ=0.3, block_shift=0, seed=7)
        goals = scen[:, 5:7].to(torch.int32).contiguous()
        ws = ffmp.ops.flow_field_workspace(n, 512, dev)
        bufs = (torch.empty((n, 512, 512), dtype=torch.int32, device=dev), torch.empty((n, 512, 512), dtype=torch.uint8, device=dev))
        for _ in range(2):
            ffmp.ops.flow_field(occ, goals, out=bufs, workspace=ws)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x.record(); ffmp.ops.flow_field(occ, goals, out=bufs, workspace=ws); y.record(); torch.cuda.synchronize()
            ts.append(x.elapsed_time(y))
        ts.sort()
        out[f"ms_{n}_grids"] = round(ts[2], 3)
        del occ, scen, bufs, ws
    env = ffmp.FFMPVectorEnv(512, grid=512, window=100, slots=3, p_occ=0.3, block_shift=0, seed=1234)
    env.reset()
    acts = torch.randint(0, 28, (60, 512), device=dev)
    env.rollout(acts[:10]); env.join(); torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record(); env.rollout(acts); env.join(); y.record(); torch.cuda.synchronize()
    out["config4_us_per_step_512_envs"] = round(x.elapsed_time(y) * 1e3 / 60, 1)
    env.close()
    env = ffmp.FFMPVectorEnv(64, grid=512, window=100, slots=3, p_occ=0.3, block_shift=0, seed=1234)
    env.reset()
    env.rollout(acts[:10, :64].contiguous()); env.join(); torch.cuda.synchronize()
    x.record(); env.rollout(acts[:, :64].contiguous()); env.join(); y.record(); torch.cuda.synchronize()
    out["config4_us_per_step_64_envs"] = round(x.elapsed_time(y) * 1e3 / 60, 1)
    out["err"] = env.error_word()
    print(json.dumps(out))
else:
    for c in ("0", "1", "0", "1"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, FFMP_FLOW_CLUSTER=c), capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-600:], flush=True)
