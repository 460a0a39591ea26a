"""Generate tests/golden/gradtts_config1_golden.pt: BASELINE config 1 in its end-to-end form on the UNMODIFIED
reference (container only) - `GradTTS(149,1,64,192,768,256,2,6,3,0.1,4,80,64,0.05,20.0,1000)`, 221 synthetic token ids,
`forward(x, x_lengths, n_timesteps=10, temperature=1.5, length_scale=0.91)` on the CPU.

The text encoder keeps its seeded random initialisation (it is outside the path: its OUTPUTS mu_x / logw / x_mask are stored as
the fixture's inputs, 75 KB); the decoder is loaded (strict) with the synthetic weights every other test uses, so nothing but
seeds and outputs needs storing.  Stored: encoder outputs, the reference's three return values (attn as one token per frame)
and the seeds.  Asserts that the oracle chain prior_expand -> reverse_diffusion reproduces the reference's decoder output.

    python scripts/make_golden_config1.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import gradtts_oracle as O  # noqa: E402
from speech_backbones_b200 import UNetConfig, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.gradtts import reference_order_noise  # noqa: E402
from _ref_import import import_gradtts  # noqa: E402

SEED, NOISE_SEED, N, TEMP, LS = 1234, 21, 10, 1.5, 0.91


def main():
    import_gradtts()
    from model import GradTTS
    torch.manual_seed(SEED)
    model = GradTTS(149, 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000).eval()
    cfg = UNetConfig()
    sd = synthetic_state_dict(cfg, SEED)
    model.decoder.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(SEED)
    x = torch.randint(0, 148, (1, 221), generator=g)
    enc = {}
    real = model.encoder.forward

    def hook(*a, **k):
        enc["out"] = real(*a, **k)
        return enc["out"]
    model.encoder.forward = hook
    torch.manual_seed(NOISE_SEED)
    y_enc, y_dec, attn = model(x, torch.tensor([221]), n_timesteps=N, temperature=TEMP, length_scale=LS)
    mu_x, logw, x_mask = (t.detach().clone() for t in enc["out"])
    Ty_ = y_dec.shape[-1] + (-y_dec.shape[-1]) % 4
    torch.manual_seed(NOISE_SEED)
    noise_tf = reference_order_noise(1, 80, Ty_, torch.float32, "cpu")
    o = O.prior_expand(mu_x, logw, x_mask, LS, TEMP, noise_tf)
    assert o["y_max_length"] == y_dec.shape[-1] and torch.equal(o["mu_y"][:, :, :o["y_max_length"]], y_enc)
    with torch.no_grad():
        y = O.reverse_diffusion(sd, cfg, o["z"], o["y_mask"], o["mu_y"], N)[:, :, :o["y_max_length"]]
    err = ((y - y_dec).norm() / y_dec.norm()).item()
    assert err < 1e-5, err
    full = attn[0, 0]                                              # [Tx, y_max]  (the reference slices the token axis: no-op here)
    tok = torch.where(full.sum(0) > 0, full.argmax(0), torch.full((full.shape[1],), -1)).to(torch.int16)
    out = dict(seed=SEED, noise_seed=NOISE_SEED, N=N, temperature=TEMP, length_scale=LS, torch=torch.__version__,
               mu_x=mu_x, logw=logw, x_mask=x_mask, y_enc=y_enc.clone(), y_dec=y_dec.clone(), tok=tok, Ty=Ty_)
    path = os.path.join(ROOT, "tests", "golden", "gradtts_config1_golden.pt")
    torch.save(out, path)
    print(f"y_max_length={y_dec.shape[-1]} (padded {Ty_}); oracle chain vs reference rel-L2 {err:.2e}; wrote {path} {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
