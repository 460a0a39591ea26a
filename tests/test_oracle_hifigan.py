"""CPU: the HiFi-GAN V1 generator oracle (the step after the path) vs the committed outputs of the
UNMODIFIED reference generator (tests/golden/hifigan_golden.pt, scripts/make_golden_hifigan.py).  Groundwork for the next
round: no product kernel exists for this row yet, so there is no GPU test."""
import os

import pytest
import torch

from oracle import hifigan_oracle as H
from speech_backbones_b200.spec import hifigan_param_spec, synthetic_hifigan_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hg_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "hifigan_golden.pt"), weights_only=False)


def test_parameter_inventory_and_known_answers(hg_golden):
    spec = hifigan_param_spec()
    assert dict(spec) == dict(H.param_spec())
    n = sum(int(torch.tensor(s).prod()) for _, s in spec)
    assert n == hg_golden["nparams"] == 13926017          # HiFi-GAN V1 generator (jik876/hifi-gan: 13.92 M)
    assert H.macs_per_mel_frame() == hg_golden["macs_per_mel_frame"]


@pytest.mark.parametrize("idx", range(3))
def test_oracle_matches_reference_golden(hg_golden, idx):
    c = hg_golden["cases"][idx]
    sd = synthetic_hifigan_state_dict(hg_golden["seed"])
    g = torch.Generator().manual_seed(hg_golden["seed"] + c["T"])
    mel = torch.randn(c["B"], 80, c["T"], generator=g)
    with torch.no_grad():
        y = H.generator(sd, mel)
    assert y.shape == (c["B"], 1, c["T"] * 256)
    assert torch.allclose(y, c["out"], rtol=1e-5, atol=1e-6)      # same build + seeds => bit-exact; slack for BLAS threads
