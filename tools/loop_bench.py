"""BASELINE config 5's data plane as a closed loop: sharded envs -> learner feed (one push kernel over NVLink peer memory) ->
a policy on the learner rank -> actions back to the shards (one small NCCL broadcast) -> next step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tools/loop_bench.py [--envs 4096] [--steps 300]

The policy is a stand-in for the reference's Q-network (train.py:231-303, out of scope): it reads the gathered observation
of every env — the flow direction under the robot from the newest local map and the bearing of the goal — and picks the
full-speed action whose turn rate follows the flow field (a data-dependent consumer of the whole feed, so nothing can be
skipped).  Prints one JSON line on rank 0: closed-loop env-steps/s, the episode statistics (all-reduce of five numbers) and,
for comparison, the same envs under uniform random actions."""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402
from flow_field_based_motion_planner_b200 import sharding  # noqa: E402

# flow direction code k (SPEC.md §4: E, NE, N, NW, W, SW, S, SE in grid axes i = x, j = y) -> heading angle
DIR_ANGLE = [0.0, math.pi / 4, math.pi / 2, 3 * math.pi / 4, math.pi, -3 * math.pi / 4, -math.pi / 2, -math.pi / 4]


def policy(obs, yaw):
    """obs: gathered views [world, N, ...]; yaw f32[world, N].  -> action ids int64[world, N]."""
    lm = obs["local_map"]                                   # u8 [world, N, 2, W, W]; robot at (W/2, W/2) of the newest frame
    w = lm.shape[-1]
    code = (lm[:, :, 1, w // 2, w // 2] // 28).long()       # 0..7 direction, 8 none
    ang = torch.tensor(DIR_ANGLE + [0.0], device=lm.device)[code.clamp(max=8)]
    # where the flow field gives no direction (goal cell / unreachable) fall back to the goal bearing
    want = torch.where(code < 8, ang - yaw, obs["relative_goal"][..., 1])
    want = torch.atan2(torch.sin(want), torch.cos(want))
    iw = torch.clamp(torch.round(want / 0.1 / 0.2).long() + 3, 0, 6)      # turn rate that closes the heading error in one 0.1 s step
    iv = torch.where(want.abs() < 0.6, torch.full_like(iw, 2), torch.zeros_like(iw))   # drive at 0.4 m/s once roughly aligned
    return 7 * iv + iw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = a.envs
    res = {"n_gpus": world, "envs_per_gpu": N}

    def run(mode):
        env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=99, env_id_base=rank * N, device=f"cuda:{local}")
        env.reset()
        feed = sharding.LearnerFeed(env, dest=[0])
        stats = sharding.EpisodeStats()
        actions_all = torch.full((world, N), 3, dtype=torch.int64, device=dev)
        yaw_all = torch.zeros((world, N), dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(rank)

        def step_loop(steps, record):
            for _ in range(steps):
                _, _, done, info = env.step(actions_all[rank])
                if record:
                    stats.update(done, info["flags"], info["episode_return"], info["episode_length"])
                feed.push()
                if world > 1:                       # the policy needs the heading, which is not part of the transition block
                    dist.all_gather_into_tensor(yaw_all.view(-1), env.pose()[:, 2].contiguous())
                else:
                    yaw_all[0].copy_(env.pose()[:, 2])
                if rank == 0:
                    obs, _, _ = feed.wait()
                    if mode == "policy":
                        actions_all.copy_(policy(obs, yaw_all))
                    else:
                        actions_all.copy_(torch.randint(0, 28, (world, N), generator=gen, device=dev))
                    feed.release()
                if world > 1:
                    dist.broadcast(actions_all, src=0)

        step_loop(20, False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_loop(a.steps, True)
        env.join()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        summary = sharding.EpisodeStats.summary(stats.all_reduce())
        err = feed.error_word()
        feed.close()
        env.close()
        return {"ms_per_step": float(t) / a.steps, "env_steps_per_s": world * N * a.steps / (float(t) * 1e-3),
                "feed_error_word": err, **summary}

    res["flow_following_policy"] = run("policy")
    res["random_actions"] = run("random")
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
