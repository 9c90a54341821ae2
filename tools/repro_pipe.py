import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flow_field_based_motion_planner_b200 as ffmp
import oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
G = int(sys.argv[3]) if len(sys.argv) > 3 else 64
W = int(sys.argv[4]) if len(sys.argv) > 4 else 32
env = ffmp.FFMPVectorEnv(N, grid=G, window=W, seed=5, device="cuda:0")
orc = oracle.OracleVectorEnv(N, grid=G, window=W, seed=5)
obs = env.reset(); orc.reset()
torch.cuda.synchronize(); print("reset ok", flush=True)
rng = np.random.default_rng(0)
for t in range(T):
    a = rng.integers(0, 28, N)
    obs, reward, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
    torch.cuda.synchronize()
    orc.step(a)
    ok = (np.array_equal(reward.cpu().numpy().view(np.uint32), orc.reward.view(np.uint32)) and
          np.array_equal(obs["local_map"].cpu().numpy(), orc.local_map) and np.array_equal(done.cpu().numpy().astype(np.uint8), orc.done))
    print("step", t, "ok" if ok else "MISMATCH", flush=True)
    if not ok:
        r_bad = np.nonzero(reward.cpu().numpy().view(np.uint32) != orc.reward.view(np.uint32))[0]
        d_bad = np.nonzero(done.cpu().numpy().astype(np.uint8) != orc.done)[0]
        lm = obs["local_map"].cpu().numpy()
        m_bad = np.nonzero((lm != orc.local_map).reshape(N, -1).any(axis=1))[0]
        print(" reward bad", r_bad[:20], len(r_bad), " done bad", d_bad[:20], len(d_bad), " map bad", m_bad[:20], len(m_bad))
        if len(m_bad):
            e = m_bad[0]
            diff = (lm[e] != orc.local_map[e])
            print("  env", e, "frames differing", diff.reshape(2, -1).sum(axis=1), "rows", np.nonzero(diff[1].any(axis=1))[0][:10], "j =", e // 296, "cta", e % 296)
        break
print("done")
