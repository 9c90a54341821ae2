"""ncu driver: the large-map flow-field operator on N generated 512x512 maps (p = 0.3)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
dev = torch.device("cuda:0")
gids = torch.arange(N, device=dev)
occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), 512, p_occ=0.3, block_shift=0, seed=7)
goals = scen[:, 5:7].contiguous()
for _ in range(2):
    cost, flow = ffmp.ops.flow_field(occ, goals)
torch.cuda.synchronize()
print("done", int(cost[0].min()))
