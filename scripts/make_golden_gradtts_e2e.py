"""Generate tests/golden/gradtts_e2e_golden.pt: the COMPLETE `GradTTS.forward` (Grad-TTS/model/tts.py:52-99) of the UNMODIFIED
reference on the CPU with every weight seeded - text encoder (oracle/text_encoder_oracle.py:synthetic_weights) and decoder
(speech_backbones_b200.spec.synthetic_state_dict) strict-loaded - so that the whole call can be rebuilt from seeds:
token ids -> TextEncoder -> durations / alignment / prior / terminal sample -> N-step reverse diffusion -> mel.
Asserts that the chain of the three oracles (text_encoder -> prior_expand -> reverse_diffusion) reproduces the reference.

    python scripts/make_golden_gradtts_e2e.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import gradtts_oracle as O  # noqa: E402
from oracle import text_encoder_oracle as T  # noqa: E402
from speech_backbones_b200 import UNetConfig, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.gradtts import reference_order_noise  # noqa: E402
from _ref_import import import_gradtts  # noqa: E402

SEED, NOISE_SEED = 2024, 5
CASES = [dict(B=2, Tx=30, lengths=[30, 17], N=4, temperature=1.5, length_scale=0.91),
         dict(B=1, Tx=64, lengths=[64], N=10, temperature=1.0, length_scale=1.0)]


def main():
    import_gradtts()
    from model import GradTTS
    model = GradTTS(149, 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000).eval()
    cfg = UNetConfig()
    sd_enc, sd_dec = T.synthetic_weights(SEED), synthetic_state_dict(cfg, SEED)
    model.encoder.load_state_dict(sd_enc, strict=True)
    model.decoder.load_state_dict(sd_dec, strict=True)
    assert model.nparams == 14835032                                  # SURVEY 8c anchor
    out = {"seed": SEED, "noise_seed": NOISE_SEED, "torch": torch.__version__, "cases": []}
    for c in CASES:
        g = torch.Generator().manual_seed(SEED + c["Tx"])
        x = torch.randint(0, 148, (c["B"], c["Tx"]), generator=g)
        xl = torch.tensor(c["lengths"])
        torch.manual_seed(NOISE_SEED)
        y_enc, y_dec, attn = model(x, xl, n_timesteps=c["N"], temperature=c["temperature"], length_scale=c["length_scale"])
        with torch.no_grad():
            mu_x, logw, x_mask = T.text_encoder(sd_enc, x, xl)
            Ty = y_dec.shape[-1] + (-y_dec.shape[-1]) % 4
            torch.manual_seed(NOISE_SEED)
            o = O.prior_expand(mu_x, logw, x_mask, c["length_scale"], c["temperature"],
                               reference_order_noise(c["B"], 80, Ty, torch.float32, "cpu"))
            L = o["y_max_length"]
            y = O.reverse_diffusion(sd_dec, cfg, o["z"], o["y_mask"], o["mu_y"], c["N"])[:, :, :L]
        assert L == y_dec.shape[-1]
        e_enc = ((o["mu_y"][:, :, :L] - y_enc).norm() / y_enc.norm()).item()
        e_dec = ((y - y_dec).norm() / y_dec.norm()).item()
        assert e_enc < 1e-5 and e_dec < 1e-4, (e_enc, e_dec)
        out["cases"].append(dict(c, y_enc=y_enc.clone(), y_dec=y_dec.clone()))
        print(f"B={c['B']} Tx={c['Tx']} N={c['N']}: frames {L}; oracle chain vs GradTTS.forward rel-L2 enc {e_enc:.1e} dec {e_dec:.1e}")
    path = os.path.join(ROOT, "tests", "golden", "gradtts_e2e_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
