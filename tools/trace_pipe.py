"""Per-env timeline of tick_pipe_kernel (FFMP_TRACE=1 FFMP_TICK_PIPE=1): cycles since CTA start."""
import ctypes as C, json, os, sys
os.environ["FFMP_TRACE"] = "1"; os.environ["FFMP_TICK_PIPE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
N = 4096
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234, p_occ=0.0, max_steps=10 ** 9)
env.reset()
acts = torch.full((20, N), 3, dtype=torch.int64, device=env.device)
env.rollout(acts); env.join(); torch.cuda.synchronize()
env.step(acts[0]); env.join(); torch.cuda.synchronize()
out = np.zeros((N, 8), dtype=np.uint64)
ffmp.native.check(env._L.ffmp_debug_trace(env._h, C.c_void_p(out.ctypes.data), env._stream()), "trace")
c = out.astype(np.int64)
j = np.arange(N) // 296          # position of the env inside its CTA
res = {}
for r in range(3):
    m = (j // 6) == r
    res[f"round{r}"] = {"p1_done": float(c[m, 0].mean()), "tma_issue": float(c[m, 1].mean()), "consumer_wait_start": float(c[m, 2].mean()),
                        "full": float(c[m, 3].mean()), "verdict": float(c[m, 4].mean()), "drained": float(c[m, 5].mean())}
print(json.dumps(res, indent=1))
