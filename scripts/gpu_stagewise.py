"""GPU debugging aid: print per-stage rel-L2 of libsbk intermediates vs the oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.binding import Engine  # noqa: E402
from test_parity_gpu import stagewise_errors  # noqa: E402

B, T, n_spks = (int(a) for a in (sys.argv[1:4] + ["2", "32", "1"][len(sys.argv[1:4]):]))
precision = sys.argv[4] if len(sys.argv) > 4 else "fp32"
cfg = UNetConfig(n_spks=n_spks)
sd = synthetic_state_dict(cfg)
eng = Engine(n_spks=n_spks, precision=precision)
eng.load_state_dict(sd)
z, mask, mu, spk, _ = synthetic_inputs(B, T, ragged=True, n_spks=n_spks)
t = torch.linspace(0.9, 0.2, B)
for n, e, m in stagewise_errors(eng, cfg, sd, z * mask, mask, mu, t, spk, masked_storage=precision != "fp32"):
    print(f"{n:48s} rel_l2={e:.3e} |ref|max={m:.3g}")
