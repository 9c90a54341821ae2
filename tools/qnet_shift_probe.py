"""Experiment: can a tcgen05 A operand start at a 128-byte row offset that is not a multiple of the 1024-byte swizzle atom?
FC layers with the TMA box loaded `shift` rows early and the descriptor started `shift` rows late; batch 100 keeps the valid rows
inside the box.  Prints the max error against the unshifted run for base_offset on / off."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import flow_field_based_motion_planner_b200 as ffmp
    from qnet_ref import seeded_case
    net, m, g, v, t = seeded_case(100)
    dev = torch.device("cuda:0")
    qn = ffmp.QNetwork(max_batch=128).load_state_dict(net.state_dict())
    q = qn(m.to(dev).bfloat16(), g.to(dev), v.to(dev), t.to(dev))
    torch.cuda.synchronize()
    torch.save(q.cpu(), sys.argv[2])
else:
    import torch
    outs = {}
    for shift, bo in ((0, 0), (1, 1), (3, 1), (8, 1), (3, 0), (1, 0), (8, 0)):
        path = f"/tmp/q_{shift}_{bo}.pt"
        env = dict(os.environ, QNET_DBG_SHIFT=str(shift), QNET_DBG_BO=str(bo))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", path], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(shift, bo, "FAILED", r.stderr[-300:])
            continue
        outs[(shift, bo)] = torch.load(path)
        print("shift", shift, "base_offset", bo, "max |q - q_unshifted| =", float((outs[(shift, bo)] - outs[(0, 0)]).abs().max()), flush=True)
