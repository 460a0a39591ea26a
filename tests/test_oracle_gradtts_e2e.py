"""CPU: the complete `GradTTS.forward` (Grad-TTS/model/tts.py:52-99) restated as a chain of the three pinned oracles -
text_encoder -> prior_expand -> reverse_diffusion - vs the committed outputs of the UNMODIFIED reference with every weight
seeded (scripts/make_golden_gradtts_e2e.py).  Token ids in, mel out: the oracle of the whole call inference.py:76 makes."""
import os

import pytest
import torch

from oracle import gradtts_oracle as O
from oracle import text_encoder_oracle as T
from speech_backbones_b200 import UNetConfig, synthetic_state_dict
from speech_backbones_b200.gradtts import reference_order_noise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("idx", range(2))
def test_oracle_chain_reproduces_gradtts_forward(idx):
    g = torch.load(os.path.join(ROOT, "tests", "golden", "gradtts_e2e_golden.pt"), weights_only=False)
    c = g["cases"][idx]
    cfg = UNetConfig()
    sd_enc, sd_dec = T.synthetic_weights(g["seed"]), synthetic_state_dict(cfg, g["seed"])
    gen = torch.Generator().manual_seed(g["seed"] + c["Tx"])
    x = torch.randint(0, 148, (c["B"], c["Tx"]), generator=gen)
    with torch.no_grad():
        mu_x, logw, x_mask = T.text_encoder(sd_enc, x, torch.tensor(c["lengths"]))
        L = c["y_dec"].shape[-1]
        torch.manual_seed(g["noise_seed"])
        o = O.prior_expand(mu_x, logw, x_mask, c["length_scale"], c["temperature"],
                           reference_order_noise(c["B"], 80, L + (-L) % 4, torch.float32, "cpu"))
        assert o["y_max_length"] == L
        y = O.reverse_diffusion(sd_dec, cfg, o["z"], o["y_mask"], o["mu_y"], c["N"])[:, :, :L]
    assert torch.allclose(o["mu_y"][:, :, :L], c["y_enc"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(y, c["y_dec"], rtol=1e-4, atol=1e-4 * c["y_dec"].abs().max().item())
