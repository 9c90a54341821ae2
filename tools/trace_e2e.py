"""Kernel span of the tick (FFMP_TRACE=1, per-CTA globaltimer stamps) in rollout mode and in host-buffer (e2e) mode."""
import ctypes as C, json, os, sys
os.environ["FFMP_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
N = 4096
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234)
env.reset()
dev = env.device
acts = torch.randint(0, 28, (50, N), device=dev)
host = [torch.randint(0, 28, (N,), dtype=torch.int64).pin_memory() for _ in range(8)]
out = np.zeros((N, 8), dtype=np.uint64)

def span():
    ffmp.native.check(env._L.ffmp_debug_trace(env._h, C.c_void_p(out.ctypes.data), env._stream()), "trace")
    g0, g1 = out[:, 0].astype(np.int64), out[:, 7].astype(np.int64)
    return {"kernel_span_us": float(g1.max() - g0.min()) / 1e3, "cta_lifetime_us_mean": float((g1 - g0).mean()) / 1e3,
            "start_spread_us_p50_p90": [float(np.percentile(g0 - g0.min(), q)) / 1e3 for q in (50, 90)]}

res = {}
spans = []
for rep in range(5):
    env.rollout(acts)                     # 50 back-to-back ticks, regeneration fully overlapped
    torch.cuda.synchronize()
    spans.append(span())
res["rollout_mode_last_tick"] = spans[-1]; res["rollout_spans"] = [s["kernel_span_us"] for s in spans]
spans = []
for rep in range(5):
    for i in range(60):
        env.step_host(host[i % 8])
    spans.append(span())
res["e2e_mode_last_tick"] = spans[-1]; res["e2e_spans"] = [s["kernel_span_us"] for s in spans]
print(json.dumps(res))
