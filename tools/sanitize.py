"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
from flow_field_based_motion_planner_b200 import sharding
dev = torch.device("cuda:0")
for grid, window, n in ((128, 100, 24), (64, 32, 16), (100, 100, 4), (256, 100, 3)):
    env = ffmp.FFMPVectorEnv(n, grid=grid, window=window, seed=3, ring=3, slots=3, max_steps=6)
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(0)
    for t in range(14):
        env.step(torch.randint(0, 28, (n,), device=dev, generator=g))
    a = torch.randint(0, 28, (n,)).pin_memory()
    env.step_host(a)
    env.scan(90, 3.5)
    env.learner_input(torch.bfloat16)
    ring = ffmp.ReplayRing(env, 3); ring.push(); ring.push(torch.zeros(n, dtype=torch.int64, device=dev)); ring.sample(8)
    feed = sharding.LearnerFeed(env); feed.push(); feed.wait(); feed.release(); feed.close()
    env.reset(torch.ones(n, dtype=torch.uint8, device=dev))
    env.join()
    torch.cuda.synchronize()
    env.close()
gids = torch.arange(6, device=dev)
for G in (32, 96, 128, 160, 512):
    occ, scen = ffmp.ops.generate_scenarios(gids, torch.zeros_like(gids), G, seed=1)
    ffmp.ops.flow_field(occ, scen[:, 5:7].contiguous())
    ffmp.ops.scan(occ, torch.tensor([[1.0, 1.0, 0.3]] * 6, device=dev), beams=33, range_max=2.0, flow_mode=False)
torch.cuda.synchronize()
print("sanitize pass done")
