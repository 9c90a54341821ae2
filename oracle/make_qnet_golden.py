"""Pin tests/qnet_ref.py to the UNMODIFIED reference `Network` and commit that class's outputs (build container only).

The class is exec'd from /root/reference/src/train.py:231-303 with the module globals it reads (BATCH_SIZE, device = cpu).
    python oracle/make_qnet_golden.py
"""
import json
import os
import sys
import textwrap

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from qnet_ref import RefNetwork, seeded_case  # noqa: E402

REF = "/root/reference/src/train.py"
OUT = os.path.join(ROOT, "tests", "golden", "qnet_golden.json")


def main():
    with open(REF) as f:
        src = textwrap.dedent("".join(f.readlines()[230:303]))
    env = {"nn": nn, "F": F, "torch": torch, "BATCH_SIZE": 1024, "device": torch.device("cpu")}
    exec(compile(src, REF + ":231-303", "exec"), env)
    out = {"seed": 1234, "cases": []}
    for batch in (1, 3):
        mine, m, g, v, t = seeded_case(batch)
        torch.manual_seed(1234)
        ref = env["Network"](2, 28).float().eval()
        for (ka, a), (kb, b) in zip(mine.state_dict().items(), ref.state_dict().items()):
            assert ka == kb and torch.equal(a, b), ka            # same seed, same creation order -> same weights
        with torch.no_grad():
            q_ref = ref(m, g, v, t)
            taps = []
            q_mine = mine(m, g, v, t, taps=taps)
        err = float((q_ref - q_mine).abs().max())
        assert err <= 1e-4 * float(q_ref.abs().max()) + 1e-5, err
        out["cases"].append({"batch": batch, "q": q_ref.tolist(), "restatement_max_abs_err": err,
                             "tap_means": [float(x.mean()) for x in taps], "tap_abs_max": [float(x.abs().max()) for x in taps]})
        print("batch", batch, "restatement vs reference max abs err", err, "q range", float(q_ref.min()), float(q_ref.max()))
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
