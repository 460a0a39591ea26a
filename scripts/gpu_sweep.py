"""mel-frames/s of the sampler at the BASELINE.json configs' shapes and N in {10, 50, 1000} (one GPU).
Writes one JSON line per case; CUDA-event timed, 1 warm-up + `reps` timed calls."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.binding import Engine  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
CASES = [  # (label, B, T, N, reps)
    ("config1 shape: single utterance, N=10", 1, 512, 10, 5),
    ("single utterance, N=50", 1, 512, 50, 3),
    ("config2: B=32 T=512 N=10", 32, 512, 10, 3),
    ("config2: B=32 T=512 N=50", 32, 512, 50, 3),
    ("config3 shape: B=128 T=512 N=50 of 1000", 128, 512, 50, 2),
    ("long horizon: B=32 T=512 N=1000", 32, 512, 1000, 1),
    ("config5 per-GPU share: B=256 T=512 N=50", 256, 512, 50, 1),
]
cfg = UNetConfig()
eng = Engine(precision=precision)
eng.load_state_dict(synthetic_state_dict(cfg))
for label, B, T, N, reps in CASES:
    z, mask, mu, _, _ = synthetic_inputs(B, T)
    zd, md, mud = z.cuda(), mask.cuda(), mu.cuda()
    eng.reverse_diffusion(zd, md, mud, min(N, 3))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = eng.reverse_diffusion(zd, md, mud, N)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"case": label, "precision": precision, "B": B, "T": T, "N": N, "ms_per_call": ms,
                      "mel_frames_per_s": B * T / (ms * 1e-3), "frame_steps_per_s": B * T * N / (ms * 1e-3),
                      "ms_per_sampler_step": ms / N, "finite": bool(torch.isfinite(y).all()),
                      "workspace_GB": eng.workspace_bytes(B, T) / 1e9}), flush=True)
    del zd, md, mud, y
