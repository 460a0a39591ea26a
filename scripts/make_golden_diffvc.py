"""Generate tests/golden/diffvc_golden.pt from the UNMODIFIED DiffVC reference (container only) and pin
oracle/diffvc_oracle.py against it.  Run in its own process (Grad-TTS and DiffVC both name their package `model`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from speech_backbones_b200.spec import DiffVCConfig, diffvc_param_spec, synthetic_diffvc_inputs, synthetic_state_dict  # noqa: E402
from _ref_import import import_diffvc  # noqa: E402

CASES = [
    dict(kind="est", B=2, T=32, Tr=24, ragged=True, t=[0.9, 0.4]),
    dict(kind="est", B=1, T=64, Tr=64, ragged=False, t=[0.05]),
    dict(kind="traj", B=2, T=32, Tr=24, ragged=True, N=6, mode="ml"),
    dict(kind="traj", B=2, T=32, Tr=24, ragged=True, N=4, mode="em"),
    dict(kind="traj", B=2, T=32, Tr=24, ragged=True, N=4, mode="pf"),
    dict(kind="traj", B=1, T=48, Tr=40, ragged=False, N=30, mode="ml"),
]
SEED, NOISE_SEED = 1234, 11


def main():
    md = import_diffvc()
    from oracle import diffvc_oracle as O
    cfg = DiffVCConfig()
    spec = diffvc_param_spec(cfg)
    ref = md.Diffusion(cfg.n_feats, cfg.dim_unet, cfg.dim_spk, cfg.use_ref_t, cfg.beta_min, cfg.beta_max).eval()
    assert sum(v.numel() for v in ref.state_dict().values()) == 117_794_599          # SURVEY.md 8(c) anchor
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v) for k, v in spec.items()}
    sd = synthetic_state_dict(cfg, SEED, spec=spec)
    ref.load_state_dict(sd, strict=True)
    out = {"seed": SEED, "noise_seed": NOISE_SEED, "cases": [], "torch": torch.__version__}
    for case in CASES:
        z, mask, mean, r, rmask, mean_ref, c = synthetic_diffvc_inputs(case["B"], case["T"], case["Tr"], seed=SEED, ragged=case["ragged"])
        with torch.no_grad():
            if case["kind"] == "est":
                t = torch.tensor(case["t"])
                xt_ref = ref.compute_diffused_mean(r, rmask, mean_ref, 0.5)[:, None]
                y_ref = ref.estimator(z * mask, mask, mean, xt_ref, rmask, c, t)
                y_orc = O.estimator(sd, cfg, z * mask, mask, mean, xt_ref, rmask, c, t)
            else:
                torch.manual_seed(NOISE_SEED)
                y_ref = ref(z, mask, mean, r, rmask, mean_ref, c, case["N"], case["mode"])
                torch.manual_seed(NOISE_SEED)
                y_orc = O.reverse_diffusion(sd, cfg, z, mask, mean, r, rmask, mean_ref, c, case["N"], case["mode"])
        err = (y_ref - y_orc).abs().max().item()
        assert torch.isfinite(y_ref).all(), case
        assert err <= 2e-5 * max(1.0, y_ref.abs().max().item()), (case, err)
        print(f"{case}: |ref|max={y_ref.abs().max().item():.4g} oracle-vs-ref max abs err={err:.3g}")
        out["cases"].append(dict(case, out=y_ref.clone()))
    path = os.path.join(ROOT, "tests", "golden", "diffvc_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
