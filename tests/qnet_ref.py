"""fp32 PyTorch restatement of the reference trainer's `Network` (/root/reference/src/train.py:231-303) for the parity tests.

The reference class cannot be imported on the GPU box (train.py needs rospy / kornia and /root/reference does not travel), so
the tests use this restatement; oracle/make_qnet_golden.py pins it to the UNMODIFIED class (exec'd from train.py:231-303) in the
build container and commits that class's outputs as tests/golden/qnet_golden.json.  Parameter creation order equals the
reference's __init__, so the same torch seed yields the same weights.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class RefNetwork(nn.Module):
    def __init__(self, input_channels=2, outputs=28):
        super().__init__()
        self.conv1 = nn.Conv2d(input_channels, 32, kernel_size=32)      # train.py:236
        self.conv2 = nn.Conv2d(32, 64, kernel_size=32)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=8)
        self.conv4 = nn.Conv2d(64, 64, kernel_size=8)
        self.fc1 = nn.Linear(5, 67)
        self.fc2 = nn.Linear(6400, 512)
        self.fc3 = nn.Linear(512, 512)
        self.fc4_ea = nn.Linear(512, outputs)
        self.fc4_ev = nn.Linear(512, 1)

    def forward(self, state_m, state_g, state_v, state_t, scalar_tile=True, taps=None):
        x = F.relu(self.conv1(state_m))
        if taps is not None: taps.append(x)
        x = F.relu(self.conv2(x))
        if taps is not None: taps.append(x)
        x = F.relu(self.conv3(x))
        # train.py:259-276: only element [0][30] of relu(fc1(cat(g, v, t))) survives, as a scalar added to every feature
        s = F.relu(self.fc1(torch.cat((state_g, state_v, state_t), 1)))[0][30].detach() if scalar_tile else 0.0
        x = x + s
        if taps is not None: taps.append(x)
        for _ in range(3):                                                # train.py:278-280: conv4 three times
            x = F.relu(self.conv4(x))
            if taps is not None: taps.append(x)
        x = torch.flatten(x, start_dim=1)
        x = F.relu(self.fc2(x))
        if taps is not None: taps.append(x)
        x = F.relu(self.fc3(x))
        if taps is not None: taps.append(x)
        adv, val = self.fc4_ea(x), self.fc4_ev(x)
        return adv + val - adv.mean(1, keepdim=True).expand(-1, adv.size(1))   # train.py:299


def seeded_case(batch, seed=1234):
    """The network and inputs of the golden vectors: weights from torch.manual_seed(seed) in the reference's creation order,
    flow-image-like maps (values 0, 28, ..., 224, 255), goal / velocity / dt in their observation ranges."""
    torch.manual_seed(seed)
    net = RefNetwork().float().eval()
    g = torch.Generator().manual_seed(seed + 1)
    codes = torch.tensor([0, 28, 56, 84, 112, 140, 168, 196, 224, 255], dtype=torch.float32)
    m = codes[torch.randint(0, 10, (batch, 2, 100, 100), generator=g)]
    sg = torch.rand((batch, 2), generator=g) * torch.tensor([7.0, 6.28]) - torch.tensor([0.0, 3.14])
    sv = torch.rand((batch, 2), generator=g) * 0.06
    st = torch.full((batch, 1), 0.1)
    return net, m, sg, sv, st
