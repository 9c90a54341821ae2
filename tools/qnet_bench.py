"""Q network forward (csrc/qnet.cu) on B200: ms and TFLOP/s per batch, against torch (cuDNN) running the fp32 restatement in
bf16 / channels_last and fp32 on the same weights.   python tools/qnet_bench.py [batch ...]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import flow_field_based_motion_planner_b200 as ffmp
from qnet_ref import seeded_case

# MACs per sample (train.py:236-242 on 100 x 100 maps): conv1 69^2*32*2048, conv2 38^2*64*32768, conv3 31^2*64*4096,
# conv4 (24^2 + 17^2 + 10^2)*64*4096, fc2 6400*512, fc3 512*512, heads 512*29
MACS = 69 * 69 * 32 * 2048 + 38 * 38 * 64 * 32768 + 31 * 31 * 64 * 4096 + (24 * 24 + 17 * 17 + 100) * 64 * 4096 + 6400 * 512 + 512 * 512 + 512 * 29


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    batches = [int(x) for x in sys.argv[1:]] or [1, 64, 1024]
    net, *_ = seeded_case(1)
    net = net.to(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    for B in batches:
        qn = ffmp.QNetwork(max_batch=B).load_state_dict(net.state_dict())
        m = (torch.randint(0, 10, (B, 2, 100, 100), device=dev) * 28).to(torch.bfloat16)
        g = torch.rand((B, 2), device=dev); v = torch.rand((B, 2), device=dev); t = torch.full((B, 1), 0.1, device=dev)
        out = torch.empty((B, 28), device=dev)
        reps = 20 if B <= 64 else 3
        ms = timed(lambda: qn(m, g, v, t, out=out), reps)
        res = {"batch": B, "ms": ms, "tflops": 2 * MACS * B / (ms * 1e-3) / 1e12, "gmac_per_sample": MACS / 1e9}
        if "bf16_tflops_sustained" in peaks:
            res["frac_of_bf16_sustained"] = res["tflops"] / peaks["bf16_tflops_sustained"]
        with torch.no_grad():
            mf = m.float()
            if B <= 1024:
                res["torch_fp32_ms"] = timed(lambda: net(mf, g, v, t), max(1, reps // 3))
            nb = net.to(torch.bfloat16).to(memory_format=torch.channels_last)
            mb = m.contiguous(memory_format=torch.channels_last)
            gb, vb, tb = g.bfloat16(), v.bfloat16(), t.bfloat16()
            res["torch_bf16_channels_last_ms"] = timed(lambda: nb(mb, gb, vb, tb), reps)
            net.float()
        print(json.dumps(res), flush=True)
        qn.close()


if __name__ == "__main__":
    main()
