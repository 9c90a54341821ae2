"""QNetwork — the reference trainer's `Network` (/root/reference/src/train.py:231-303) evaluated by the tcgen05 implicit-GEMM
kernels of csrc/qnet.cu.  Host side only: weight hand-over and the ctypes calls; no CPU fallback.

    net = QNetwork(max_batch=4096)
    net.load_state_dict(reference_network.state_dict())          # the module's own names: conv1.weight, ..., fc4_ev.bias
    q = net(state_m, state_g, state_v, state_t)                   # f32 [B, 28], as Network.forward returns

`state_m` is [B, 2, 100, 100] (NCHW), bfloat16 (FFMPVectorEnv.learner_input(dtype=torch.bfloat16) writes exactly that) or
float32.  `scalar_tile=True` (default) keeps the reference's behaviour that only `relu(fc1(cat(g, v, t)))[0][30]` — one scalar
of sample 0 — is added to the conv3 feature map (train.py:264-276); False leaves the tile out.
"""
import ctypes as C

import torch

from . import native

LAYERS = ("conv1", "conv2", "conv3", "conv4", "fc1", "fc2", "fc3", "fc4_ea", "fc4_ev")
SHAPES = {"conv1": (32, 2, 32, 32), "conv2": (64, 32, 32, 32), "conv3": (64, 64, 8, 8), "conv4": (64, 64, 8, 8),
          "fc1": (67, 5), "fc2": (512, 6400), "fc3": (512, 512), "fc4_ea": (28, 512), "fc4_ev": (1, 512)}
NUM_ACTIONS = 28


def _check(rc, what):
    if rc != 0:
        raise native.NativeError(f"{what} failed ({rc}): {native.lib().ffmp_qnet_last_error().decode()}")


class QNetwork:
    def __init__(self, max_batch=1024, device="cuda:0", scalar_tile=True):
        self._L = native.lib()
        if not torch.cuda.is_available():
            raise native.NativeError("QNetwork needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.max_batch = int(max_batch)
        self.scalar_tile = bool(scalar_tile)
        self._h = C.c_void_p()
        _check(self._L.ffmp_qnet_create(idx, self.max_batch, C.byref(self._h)), "ffmp_qnet_create")
        self._params = None

    def _stream(self):
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))

    def load_state_dict(self, sd):
        """sd: the reference module's state_dict (any device / float dtype); converted to the GEMM layouts on the device."""
        keep = []
        w = (C.c_void_p * 9)()
        b = (C.c_void_p * 9)()
        for i, name in enumerate(LAYERS):
            wt = sd[name + ".weight"].detach().to(device=self.device, dtype=torch.float32).contiguous()
            bt = sd[name + ".bias"].detach().to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(wt.shape) != SHAPES[name] or bt.numel() != SHAPES[name][0]:
                raise ValueError(f"{name}: expected weight {SHAPES[name]}, got {tuple(wt.shape)}")
            keep += [wt, bt]
            w[i], b[i] = wt.data_ptr(), bt.data_ptr()
        with torch.cuda.device(self.device):
            _check(self._L.ffmp_qnet_load(self._h, w, b, self._stream()), "ffmp_qnet_load")
            torch.cuda.current_stream().synchronize()       # the f32 sources may be released after this
        self._params = True
        return self

    def forward(self, state_m, state_g, state_v, state_t, out=None):
        B = state_m.shape[0]
        if tuple(state_m.shape[1:]) != (2, 100, 100) or state_m.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("state_m must be [B, 2, 100, 100] float32 or bfloat16")
        m = state_m.to(self.device).contiguous()
        g = state_g.to(device=self.device, dtype=torch.float32).reshape(B, 2).contiguous()
        v = state_v.to(device=self.device, dtype=torch.float32).reshape(B, 2).contiguous()
        t = state_t.to(device=self.device, dtype=torch.float32).reshape(B, 1).contiguous()
        if out is None:
            out = torch.empty((B, NUM_ACTIONS), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self._L.ffmp_qnet_forward(self._h, B, C.c_void_p(m.data_ptr()), 1 if m.dtype == torch.bfloat16 else 0,
                                             C.c_void_p(g.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(t.data_ptr()),
                                             1 if self.scalar_tile else 0, C.c_void_p(out.data_ptr()), self._stream()),
                   "ffmp_qnet_forward")
        return out

    __call__ = forward

    def activation(self, layer, batch):
        """Intermediate activation of the last forward (tests): layers 1..6 -> bf16 [B, OH, OW, C] (NHWC), 7 / 8 -> [B, 512]."""
        per = C.c_size_t()
        _check(self._L.ffmp_qnet_debug_activation(self._h, layer, batch, None, C.byref(per), None), "ffmp_qnet_debug_activation")
        buf = torch.empty((batch, per.value), dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            _check(self._L.ffmp_qnet_debug_activation(self._h, layer, batch, C.c_void_p(buf.data_ptr()), C.byref(per), self._stream()),
                   "ffmp_qnet_debug_activation")
        return buf

    def launch_count(self):
        n = C.c_uint64()
        _check(self._L.ffmp_qnet_launch_count(self._h, C.byref(n)), "ffmp_qnet_launch_count")
        return n.value

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.ffmp_qnet_destroy(h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
