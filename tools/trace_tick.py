"""Per-CTA timeline of the tick kernel (FFMP_TRACE=1): where the latency of one env step goes."""
import ctypes as C
import json
import os
import sys

os.environ["FFMP_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
still = len(sys.argv) > 2 and sys.argv[2] == "still"
kw = dict(p_occ=0.0, max_steps=10 ** 9) if still else {}
env = ffmp.FFMPVectorEnv(N, grid=128, window=100, seed=1234, **kw)
env.reset()
dev = env.device
acts = torch.full((50, N), 3, dtype=torch.int64, device=dev) if still else torch.randint(0, 28, (50, N), device=dev)
env.rollout(acts)
env.join()
torch.cuda.synchronize()
env.step(acts[0])
env.join()
torch.cuda.synchronize()
out = np.zeros((N, 8), dtype=np.uint64)
ffmp.native.check(env._L.ffmp_debug_trace(env._h, C.c_void_p(out.ctypes.data), env._stream()), "trace")
g0, g1 = out[:, 0].astype(np.int64), out[:, 7].astype(np.int64)
c = out.astype(np.int64)
c1 = c[:, 1] >> 8
smid = c[:, 1] & 0xFF
res = {"N": N, "still": still,
       "kernel_span_us": float(g1.max() - g0.min()) / 1e3,
       "cta_lifetime_us_mean": float((g1 - g0).mean()) / 1e3, "cta_lifetime_us_p90": float(np.percentile(g1 - g0, 90)) / 1e3,
       "start_spread_us_p50_p90_max": [float(np.percentile(g0 - g0.min(), q)) / 1e3 for q in (50, 90, 100)],
       "cycles_start_to_kinematics": float((c[:, 2] - c1).mean()),
       "cycles_kinematics_to_waitstart": float((c[:, 3] - c[:, 2]).mean()),
       "cycles_window_wait": float((c[:, 4] - c[:, 3]).mean()),
       "cycles_wait_to_drainstart": float((c[:, 5] - c[:, 4]).mean()),
       "cycles_drain": float((c[:, 6] - c[:, 5]).mean()),
       "ctas_per_sm_max": int(np.bincount(smid).max()), "sms_used": int((np.bincount(smid) > 0).sum())}
print(json.dumps(res))
