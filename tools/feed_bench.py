"""Learner feed over NCCL (BASELINE config 5 shape): every rank steps its env shard and all-gathers the packed
transition block [local_map u8 N*2*W*W | relative_goal, velocity f32 | reward f32 | done u8] every step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/feed_bench.py [--envs 4096] [--steps 200]

Prints one JSON line on rank 0: env-steps/s with and without the per-step gather, the gather's own time (CUDA events, max over
ranks) and the bytes every rank receives per step."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402
from flow_field_based_motion_planner_b200 import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--window", type=int, default=100)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, W = a.envs, a.window
    env = ffmp.FFMPVectorEnv(N, grid=128, window=W, seed=1234, env_id_base=rank * N, device=f"cuda:{local}")
    env.reset()
    acts = torch.randint(0, 28, (64, N), device=dev, dtype=torch.int64)
    per = sharding.transition_nbytes(N, W)
    local_buf = torch.empty(per, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * per, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(steps, gather):
        g_ms = 0.0
        evs = []
        for t in range(steps):
            obs, reward, done, _ = env.step(acts[t % 64])
            if gather:
                x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x.record()
                sharding.pack_transitions(obs, reward, done, out=local_buf)
                if world > 1:
                    dist.all_gather_into_tensor(gathered, local_buf)
                else:
                    gathered.copy_(local_buf)
                y.record()
                evs.append((x, y))
        env.join()
        torch.cuda.synchronize()
        for x, y in evs:
            g_ms += x.elapsed_time(y)
        return g_ms / max(1, len(evs))

    res = {}
    for gather in (False, True):
        run(20, gather)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g_ms = run(a.steps, gather)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, g_ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        key = "gather_every_step" if gather else "no_gather"
        res[key] = {"ms_per_step": float(t[0]) / a.steps, "env_steps_per_s": world * N * a.steps / (float(t[0]) * 1e-3)}
        if gather:
            res[key]["pack_plus_all_gather_ms"] = float(t[1])
            res[key]["bytes_received_per_rank_per_step"] = (world - 1) * per
            res[key]["ingress_gbs"] = (world - 1) * per / (float(t[1]) * 1e-3) / 1e9 if world > 1 else None
    # ---- the same exchange as ONE push kernel over NVLink peer memory (csrc/feed.cu) ----
    for name, dest in (("peer_push_all_ranks", None), ("peer_push_to_learner_rank0", [0])):
        feed = sharding.LearnerFeed(env, dest=dest)
        consumer = dest is None or rank in dest

        def run_push(steps):
            evs = []
            for t in range(steps):
                env.step(acts[t % 64])
                x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x.record()
                feed.push()
                if consumer:
                    feed.wait()
                    feed.release()
                y.record()
                evs.append((x, y))
            env.join()
            torch.cuda.synchronize()
            return sum(x.elapsed_time(y) for x, y in evs) / max(1, len(evs))

        run_push(20)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g_ms = run_push(a.steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, g_ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # check: what arrived equals what NCCL gathers for the same step
        obs, reward, done, _ = env.step(acts[0])
        sharding.pack_transitions(obs, reward, done, out=local_buf)
        if world > 1:
            dist.all_gather_into_tensor(gathered, local_buf)
        else:
            gathered.copy_(local_buf)
        feed.push()
        same = True
        if consumer:
            feed.wait()
            torch.cuda.synchronize()
            for r in range(world):
                slot = feed._mem[(feed.seq & 1) * feed.buffer_stride + r * feed.slot_stride:][:per]
                same = same and bool(torch.equal(slot, gathered[r * per:(r + 1) * per]))
            feed.release()
        torch.cuda.synchronize()
        err = feed.error_word()
        flag = torch.tensor([1 if (same and err == 0) else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        n_dest = world if dest is None else len(dest)
        res[name] = {"ms_per_step": float(t[0]) / a.steps, "env_steps_per_s": world * N * a.steps / (float(t[0]) * 1e-3),
                     "push_wait_release_ms": float(t[1]), "equals_nccl_gather": bool(flag.item()),
                     "ingress_gbs_at_a_destination": (world - 1) * per / (float(t[1]) * 1e-3) / 1e9 if world > 1 else None,
                     "destinations": n_dest}
        barrier()
        feed.close()
        barrier()
    # parity of the exchange: every rank's slot of the gathered buffer decodes to what that rank packed
    obs_g, rew_g, done_g = sharding.unpack_transitions(gathered[rank * per:(rank + 1) * per], N, W)
    ok = bool(torch.equal(obs_g["relative_goal"], env.rel_goal) and torch.equal(rew_g, env.reward))
    if rank == 0:
        print(json.dumps({"n_gpus": world, "envs_per_gpu": N, "window": W, "transition_bytes_per_rank": per,
                          "own_slot_round_trip_ok": ok, **res}), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
