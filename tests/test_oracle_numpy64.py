"""CPU: the independent float64 numpy restatement (oracle/gradtts_numpy64.py: explicit shifted-slice convolutions, no torch
operators) vs the committed outputs of the UNMODIFIED reference.  The torch oracle is bit-identical to the reference but
shares its ATen kernels; this one shares nothing, so agreement at the fp32-vs-fp64 floor (1e-6) confirms that
the restated semantics - padding, tap order, the transposed-conv phase rule, GroupNorm, the softmax axis, the Euler step -
are the reference's."""
import pytest
import torch

from helpers import case_id, case_inputs, rel_l2
from oracle import gradtts_numpy64 as N64


@pytest.mark.parametrize("pick", [dict(kind="est", T=4), dict(kind="est", T=32, scale=1.0), dict(kind="est", T=64),
                                  dict(kind="traj", T=32, N=1)])
def test_numpy64_restatement_matches_reference_golden(golden, pick):
    c = next(c for c in golden["cases"] if c["n_spks"] == 1 and all(c.get(k) == v for k, v in pick.items()))
    cfg, sd, z, mask, mu, spk = case_inputs(golden, c)
    if c["kind"] == "est":
        y = N64.estimator(sd, cfg, z * mask * c["scale"], mask, mu, torch.tensor(c["t"]))
    else:
        y = N64.reverse_diffusion(sd, cfg, z, mask, mu, c["N"])
    err = rel_l2(torch.from_numpy(y), c["out"])
    print(case_id(c), "fp64 numpy vs fp32 reference rel_l2", err)
    assert err < 2e-5
