"""One estimator call (no graph) for ncu captures: python scripts/gpu_one_call.py B T precision"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.binding import Engine  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
precision = sys.argv[3] if len(sys.argv) > 3 else "tf32"
eng = Engine(precision=precision, use_graph=False)
eng.load_state_dict(synthetic_state_dict(UNetConfig()))
z, mask, mu, _, _ = synthetic_inputs(B, T)
zd, md, mud = z.cuda(), mask.cuda(), mu.cuda()
t = torch.full((B,), 0.5, device="cuda")
out = eng.estimator(zd, md, mud, t)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
