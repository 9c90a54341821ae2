"""DDQN learner of the reference trainer (`Brain`, /root/reference/src/train.py:306-431) over the batched CUDA env.

What runs where:
  * acting (train.py:336-347 decide_action) and the two no-grad evaluations of an update — argmax_a Q_main(s', a) and
    Q_target(s', a*) (train.py:381-397) — run on the hand-written tcgen05 kernels (QNetwork, csrc/qnet.cu);
  * the gradient step (train.py:399-417: MSE loss, Adam, lr 5e-4) differentiates a torch restatement of the same module
    (`TorchNetwork`, cuDNN through autograd: library code, bf16 autocast); its updated weights are handed back to the
    kernels (ffmp_qnet_load) after every step;
  * minibatches come from the device replay ring through one gather kernel (ReplayRing.sample_learner, csrc/replay.cu).
With torch.distributed initialised the update is data parallel: every rank samples its own shard's ring, gradients are
averaged with one NCCL all-reduce, so all ranks keep identical weights and act locally (no observation traffic).
Reference semantics kept: gamma 0.95, no terminal masking of the bootstrap (the `non_final_mask` lines are commented out,
train.py:383-397), epsilon = 0.5 / (episode + 1), fc1 never receives a gradient (`.item()`, train.py:265).
"""
import copy
import random

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .qnet import NUM_ACTIONS, QNetwork

GAMMA = 0.95               # train.py:59
LEARNING_RATE = 0.0005     # train.py:71
BATCH_SIZE = 1024          # train.py:62


class TorchNetwork(nn.Module):
    """train.py:231-303, for autograd only (parameter names and creation order as the reference's)."""

    def __init__(self, input_channels=2, outputs=NUM_ACTIONS):
        super().__init__()
        self.conv1 = nn.Conv2d(input_channels, 32, kernel_size=32)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=32)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=8)
        self.conv4 = nn.Conv2d(64, 64, kernel_size=8)
        self.fc1 = nn.Linear(5, 67)
        self.fc2 = nn.Linear(6400, 512)
        self.fc3 = nn.Linear(512, 512)
        self.fc4_ea = nn.Linear(512, outputs)
        self.fc4_ev = nn.Linear(512, 1)

    def forward(self, state_m, state_g, state_v, state_t, scalar_tile=True):
        x = F.relu(self.conv1(state_m))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        if scalar_tile:
            x = x + F.relu(self.fc1(torch.cat((state_g, state_v, state_t), 1).float()))[0][30].detach().to(x.dtype)
        for _ in range(3):
            x = F.relu(self.conv4(x))
        x = torch.flatten(x, start_dim=1)
        x = F.relu(self.fc2(x))
        x = F.relu(self.fc3(x))
        adv, val = self.fc4_ea(x), self.fc4_ev(x)
        return adv + val - adv.mean(1, keepdim=True).expand(-1, adv.size(1))


class DDQNLearner:
    def __init__(self, device="cuda:0", max_batch=BATCH_SIZE, seed=0, scalar_tile=True, dt=0.1, lr=LEARNING_RATE, gamma=GAMMA):
        self.device = torch.device(device)
        self.gamma, self.dt, self.scalar_tile = float(gamma), float(dt), bool(scalar_tile)
        torch.manual_seed(seed)
        self.module = TorchNetwork().to(self.device).to(memory_format=torch.channels_last)      # f32 master weights
        self.optimizer = torch.optim.Adam(self.module.parameters(), lr=lr)
        self.main = QNetwork(max_batch=max_batch, device=str(self.device), scalar_tile=scalar_tile)
        self.target = QNetwork(max_batch=max_batch, device=str(self.device), scalar_tile=scalar_tile)
        self.main.load_state_dict(self.module.state_dict())
        self.target.load_state_dict(self.module.state_dict())
        self.loss = None
        self.updates = 0
        self._rng = random.Random(seed)

    # -- acting ------------------------------------------------------------------------------------------------------------
    def q_values(self, env, obs):
        """Q(s, .) of every env of a FFMPVectorEnv from its current observation (state_m = learner_input in bf16)."""
        x = env.learner_input(dtype=torch.bfloat16)
        t = torch.full((env.num_envs, 1), self.dt, device=self.device)
        return self.main(x, obs["relative_goal"], obs["velocity"], t)

    def act(self, env, obs, episode=0, generator=None):
        """decide_action (train.py:336-347) for every env: greedy with probability 1 - eps, eps = 0.5 / (episode + 1)."""
        q = self.q_values(env, obs)
        greedy = q.argmax(1)
        eps = 0.5 / (episode + 1)
        explore = torch.rand(env.num_envs, device=self.device, generator=generator) < eps
        rand = torch.randint(0, NUM_ACTIONS, (env.num_envs,), device=self.device, generator=generator)
        return torch.where(explore, rand, greedy)

    # -- learning ----------------------------------------------------------------------------------------------------------
    def update(self, batch, mask_terminal=False):
        """One Brain.replay step (train.py:316-333) on a minibatch from ReplayRing.sample_learner.  Returns the loss tensor."""
        B = batch["reward"].shape[0]
        t = torch.full((B, 1), self.dt, device=self.device)
        # [3-3] double-Q target, no grad, on the tcgen05 kernels
        q_next_main = self.main(batch["observe_m"], batch["observe_g"], batch["observe_v"], t)
        a_star = q_next_main.argmax(1, keepdim=True)
        next_v = self.target(batch["observe_m"], batch["observe_g"], batch["observe_v"], t).gather(1, a_star).squeeze(1)
        if mask_terminal:
            next_v = next_v * (batch["done"] == 0)
        expected = (batch["reward"] + self.gamma * next_v).detach()
        # [3-2] + [4]: Q(s, a) with grad through the torch restatement, MSE, Adam
        self.module.train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            q = self.module(batch["state_m"].contiguous(memory_format=torch.channels_last), batch["state_g"], batch["state_v"], t,
                            scalar_tile=self.scalar_tile)
        sa = q.float().gather(1, batch["action"].view(-1, 1)).squeeze(1)
        loss = F.mse_loss(sa, expected)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            grads = [p.grad for p in self.module.parameters() if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= dist.get_world_size()
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        self.optimizer.step()
        self.main.load_state_dict(self.module.state_dict())       # hand the new weights to the kernels
        self.loss = loss.detach()
        self.updates += 1
        return self.loss

    def update_target(self):
        """update_target_q_network (train.py:419-420)."""
        self.target.load_state_dict(self.module.state_dict())

    def state_dict(self):
        return {"module": copy.deepcopy(self.module.state_dict()), "optimizer": copy.deepcopy(self.optimizer.state_dict()),
                "updates": self.updates}

    def load_state_dict(self, sd):
        self.module.load_state_dict(sd["module"])
        self.optimizer.load_state_dict(sd["optimizer"])
        self.updates = sd.get("updates", 0)
        self.main.load_state_dict(self.module.state_dict())
        self.target.load_state_dict(self.module.state_dict())

    def close(self):
        self.main.close()
        self.target.close()
