"""BASELINE config 5: the closed training loop over sharded envs (the reference's main(), /root/reference/src/train.py:494-699,
with its DDQN Brain, train.py:306-431).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29552 \
        tools/train_loop.py [--envs 256] [--steps 40] [--update-every 4] [--batch 1024] [--feed push|nccl|none]

Every rank: FFMPVectorEnv shard -> Q values of its own envs on the tcgen05 kernels -> epsilon-greedy actions -> env.step ->
device replay ring.  Every --update-every steps a data-parallel DDQN update (ring gather kernel, double-Q targets on the
kernels, gradient through torch autograd, one NCCL all-reduce of the gradients).  --feed push / nccl additionally ships every
step's transition block to rank 0 (the north star's "all-gather of obs to the learner GPU": the peer-memory push kernel or the
NCCL all-gather), where the learner rank consumes it for the global reward / done statistics.
Prints one JSON line on rank 0: closed-loop env-steps/s, the split of the step time, loss and episode statistics."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402
from flow_field_based_motion_planner_b200 import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256, help="envs per GPU")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--update-every", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1024, help="global minibatch (train.py:62)")
    ap.add_argument("--feed", default="push", choices=["push", "nccl", "none"])
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = a.envs
    env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=7, env_id_base=rank * N, device=f"cuda:{local}")
    obs = env.reset()
    ring = ffmp.ReplayRing(env, capacity_steps=max(8, 20000 // max(1, N * world) + 2))      # CAPACITY = 20000 transitions (train.py:65)
    ring.push()
    learner = ffmp.DDQNLearner(device=f"cuda:{local}", max_batch=max(N, a.batch // world), seed=0, dt=env.config.dt)
    feed = sharding.LearnerFeed(env, dest=[0]) if a.feed == "push" else None
    stats = sharding.EpisodeStats()
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + rank)
    per_rank_batch = max(1, a.batch // world)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in ("act", "env", "feed", "update")}
    acc = {k: 0.0 for k in ev}
    seen_reward = torch.zeros((), device=dev, dtype=torch.float64)

    def loop(steps, timed):
        nonlocal obs, seen_reward
        for s in range(steps):
            ev["act"][0].record()
            actions = learner.act(env, obs, episode=learner.updates, generator=gen)
            ev["act"][1].record()
            ev["env"][0].record()
            obs, reward, done, info = env.step(actions)
            ring.push(actions)
            ev["env"][1].record()
            stats.update(done, info["flags"], info["episode_return"], info["episode_length"])
            ev["feed"][0].record()
            if a.feed == "push":
                feed.push()
                if rank == 0:
                    _, r_all, _ = feed.wait()
                    seen_reward += r_all.sum(dtype=torch.float64)
                    feed.release()
            elif a.feed == "nccl":
                _, r_all, _ = sharding.all_gather_transitions(obs, reward, done, env.config.window)
                if rank == 0:
                    seen_reward += r_all.sum(dtype=torch.float64)
            ev["feed"][1].record()
            ev["update"][0].record()
            if (s + 1) % a.update_every == 0 and len(ring) >= per_rank_batch:
                learner.update(ring.sample_learner(per_rank_batch, generator=gen))
                if learner.updates % 2 == 0:                 # UPDATE_TARGET_EPISODE = 2 (train.py:77), per update here
                    learner.update_target()
            ev["update"][1].record()
            if timed:
                torch.cuda.synchronize()
                for k in ev:
                    acc[k] += ev[k][0].elapsed_time(ev[k][1])

    loop(max(2, a.update_every), False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(a.steps, True)
    env.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    summary = sharding.EpisodeStats.summary(stats.all_reduce())
    if rank == 0:
        res = {"n_gpus": world, "envs_per_gpu": N, "steps": a.steps, "update_every": a.update_every, "global_batch": a.batch,
               "feed": a.feed, "env_steps_per_s": world * N * a.steps / float(t), "ms_per_step": float(t) / a.steps * 1e3,
               "ms_per_step_split": {k: v / a.steps for k, v in acc.items()}, "updates": learner.updates,
               "loss": float(learner.loss) if learner.loss is not None else None,
               "feed_reward_sum_seen_by_learner_rank": float(seen_reward), "qnet_launches": learner.main.launch_count(), **summary}
        print(json.dumps(res), flush=True)
    if feed is not None:
        feed.close()
    learner.close()
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
