"""CPU: the text-encoder oracle (Grad-TTS/model/text_encoder.py restated, with the windowed relative-position attention
written directly instead of through the reference's skewing tricks) vs the committed outputs of the UNMODIFIED reference
TextEncoder (scripts/make_golden_text_encoder.py).  Groundwork for SURVEY.md 8f rank 4; no product kernel, no GPU test."""
import os

import pytest
import torch

from oracle import text_encoder_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def te_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "text_encoder_golden.pt"), weights_only=False)


def test_inventory():
    n = sum(int(torch.tensor(s).prod()) for _, s in T.param_spec())
    assert n == 7_200_145          # GradTTS total 14,835,032 = this + the 7,634,887-parameter decoder


@pytest.mark.parametrize("idx", range(3))
def test_oracle_matches_reference_golden(te_golden, idx):
    c = te_golden["cases"][idx]
    sd = T.synthetic_weights(te_golden["seed"])
    g = torch.Generator().manual_seed(te_golden["seed"] + c["Tx"])
    x = torch.randint(0, 148, (c["B"], c["Tx"]), generator=g)
    with torch.no_grad():
        mu, logw, mask = T.text_encoder(sd, x, torch.tensor(c["lengths"]))
    assert mask.sum(-1).flatten().tolist() == c["lengths"]
    assert torch.allclose(mu, c["mu"], rtol=1e-4, atol=1e-5) and torch.allclose(logw, c["logw"], rtol=1e-4, atol=1e-5)
