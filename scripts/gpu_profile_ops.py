"""Per-launch CUDA-event profile of one reverse step (sbk_profile_ops): ms, TFLOP/s, GB/s per launch."""
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
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
cfg = UNetConfig()
eng = Engine(precision=precision)
eng.load_state_dict(synthetic_state_dict(cfg))
z, mask, mu, _, _ = synthetic_inputs(B, T)
zd, md, mud = z.cuda(), mask.cuda(), mu.cuda()
eng.reverse_diffusion(zd, md, mud, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.reverse_diffusion(zd, md, mud, 10)
e1.record()
torch.cuda.synchronize()
print(f"# B={B} T={T} {precision}: {e0.elapsed_time(e1) / 10:.3f} ms/step (graph replay)")
rows = eng.profile_ops()
rows = eng.profile_ops()
tot = sum(r[1] for r in rows)
print(f"# sum of per-launch event times: {tot:.3f} ms")
for n, ms, fl, by in rows:
    print(f"{n:44s} {ms:8.4f} ms  {fl / ms / 1e9 if ms else 0:8.1f} TFLOP/s  {by / ms / 1e6 if ms else 0:8.1f} GB/s")
