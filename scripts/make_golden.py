"""Generate tests/golden/gradtts_golden.pt from the UNMODIFIED reference (container only).

Runs /root/reference/Grad-TTS/model/diffusion.py (imported, not copied) on the seeded
synthetic weights/inputs of speech_backbones_b200.spec and stores ONLY the reference
outputs plus the case descriptions; tests regenerate weights/inputs from the seeds.
Also asserts that oracle/gradtts_oracle.py agrees with the reference on every case
(this is what pins the oracle) and that load_state_dict(strict=True) accepts the
synthetic state_dict (names/shapes == reference).

    python scripts/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from speech_backbones_b200 import (UNetConfig, synthetic_inputs, synthetic_state_dict)  # noqa: E402
from oracle import gradtts_oracle as O  # noqa: E402
from _ref_import import import_gradtts  # noqa: E402

CASES = [
    # estimator single calls: kind, n_spks, B, T, ragged, t-list, xt scale
    dict(kind="est", n_spks=1, B=2, T=32, ragged=True, t=[0.995, 0.5], scale=1.0),
    dict(kind="est", n_spks=1, B=1, T=64, ragged=False, t=[0.005], scale=1.0),
    dict(kind="est", n_spks=1, B=2, T=32, ragged=True, t=[0.3, 0.7], scale=100.0),
    dict(kind="est", n_spks=1, B=3, T=100, ragged=True, t=[0.9, 0.1, 0.5], scale=1.0),
    dict(kind="est", n_spks=1, B=1, T=4, ragged=False, t=[0.5], scale=1.0),
    dict(kind="est", n_spks=4, B=2, T=32, ragged=True, t=[0.6, 0.2], scale=1.0),
    dict(kind="est", n_spks=1, B=1, T=256, ragged=False, t=[0.5], scale=1.0),
    # trajectories: kind, n_spks, B, T, ragged, N, stoc
    dict(kind="traj", n_spks=1, B=2, T=32, ragged=True, N=1, stoc=False),
    dict(kind="traj", n_spks=1, B=2, T=32, ragged=True, N=10, stoc=False),
    dict(kind="traj", n_spks=1, B=2, T=32, ragged=True, N=5, stoc=True),
    dict(kind="traj", n_spks=4, B=2, T=32, ragged=True, N=3, stoc=False),
    dict(kind="traj", n_spks=1, B=1, T=128, ragged=False, N=10, stoc=False),   # config 1 shape (PR1 ref, CPU)
    dict(kind="traj", n_spks=1, B=3, T=52, ragged=True, N=50, stoc=False),
]
SEED = 1234
NOISE_SEED = 7


def main():
    md = import_gradtts()
    out = {"seed": SEED, "noise_seed": NOISE_SEED, "cases": [], "torch": torch.__version__}
    refs = {}
    for case in CASES:
        cfg = UNetConfig(n_spks=case["n_spks"])
        sd = synthetic_state_dict(cfg, SEED)
        if case["n_spks"] not in refs:
            m = md.Diffusion(cfg.n_feats, cfg.dim, n_spks=cfg.n_spks, spk_emb_dim=cfg.spk_emb_dim,
                             beta_min=cfg.beta_min, beta_max=cfg.beta_max, pe_scale=cfg.pe_scale).eval()
            m.load_state_dict(sd, strict=True)
            refs[case["n_spks"]] = m
        ref = refs[case["n_spks"]]
        z, mask, mu, spk, _ = synthetic_inputs(case["B"], case["T"], seed=SEED, ragged=case["ragged"],
                                               n_spks=cfg.n_spks)
        with torch.no_grad():
            if case["kind"] == "est":
                t = torch.tensor(case["t"], dtype=torch.float32)
                xt = z * mask * case["scale"]
                y_ref = ref.estimator(xt, mask, mu, t, spk)
                y_orc = O.estimator(sd, cfg, xt, mask, mu, t, spk)
            else:
                torch.manual_seed(NOISE_SEED)
                y_ref = ref(z, mask, mu, case["N"], case["stoc"], spk)
                torch.manual_seed(NOISE_SEED)
                y_orc = O.reverse_diffusion(sd, cfg, z, mask, mu, case["N"], case["stoc"], spk)
        err = (y_ref - y_orc).abs().max().item()
        assert torch.isfinite(y_ref).all(), case
        assert err <= 1e-6 * max(1.0, y_ref.abs().max().item()), (case, err)
        print(f"{case}: |ref|max={y_ref.abs().max().item():.4g} oracle-vs-ref max abs err={err:.3g}")
        out["cases"].append(dict(case, out=y_ref.clone()))
    path = os.path.join(ROOT, "tests", "golden", "gradtts_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
