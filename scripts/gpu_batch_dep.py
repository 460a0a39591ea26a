"""Batch-independence probe: run one estimator call on a batch of 3 and on its middle utterance alone and compare EVERY
named intermediate of that utterance (first stage whose in-batch / alone results differ is the culprit); also the same
batch twice (run-to-run determinism).  usage: python scripts/gpu_batch_dep.py [T=512]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from helpers import rel_l2  # noqa: E402
from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.binding import Engine  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = UNetConfig()
sd = synthetic_state_dict(cfg)
z, mask, mu, _, _ = synthetic_inputs(3, T, ragged=True)
t = torch.tensor([0.9, 0.5, 0.1])
xt = z * mask


def capture(eng, sl):
    eng.debug_capture(True)
    y = eng.estimator(xt[sl].cuda(), mask[sl].cuda(), mu[sl].cuda(), t[sl].cuda()).cpu()
    torch.cuda.synchronize()
    eng.debug_capture(False)
    out = {}
    for name in eng.debug_names():
        g = eng.debug_read(name)
        if g is not None:
            out[name] = g.clone()
    out["estimator.out"] = y.reshape(-1)
    return out


for prec in ("tf32", "bf16", "fp32"):
    eng = Engine(precision=prec)
    eng.load_state_dict(sd)
    a = capture(eng, slice(0, 3))
    a2 = capture(eng, slice(0, 3))
    b = capture(eng, slice(1, 2))
    print(f"== {prec} T={T}: same batch twice, then utterance 1 in-batch vs alone", flush=True)
    first = None
    for name in a:
        n = a[name].numel() // 3
        rr = rel_l2(a2[name], a[name])
        d = rel_l2(b[name], a[name][n:2 * n])
        flag = ""
        if first is None and d > 1e-5:
            first, flag = name, "   <== first divergent stage"
        print(f"  {name:44s} run-to-run {rr:.2e}   alone-vs-in-batch {d:.2e}{flag}")
    eng.close()
