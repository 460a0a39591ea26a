"""CPU (gloo, world_size 2): the multi-GPU host logic - weight broadcast, batch sharding, mel gather.
The compute function is injected (here: the CPU oracle), so no GPU is needed; on GPUs bench.py passes the
libsbk-backed module instead."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_backbones_b200 import UNetConfig, estimator_param_spec, synthetic_inputs, synthetic_state_dict
from speech_backbones_b200.sharded import (broadcast_state_dict, flatten_state_dict, shard_bounds, sharded_sample,
                                           unflatten_state_dict)


def test_shard_bounds_partition():
    for n in (1, 2, 5, 32, 2048):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flatten_roundtrip():
    cfg = UNetConfig()
    spec = estimator_param_spec(cfg)
    sd = synthetic_state_dict(cfg)
    flat = flatten_state_dict(sd, list(spec))
    assert flat.numel() == 7_634_887
    back = unflatten_state_dict(flat, list(spec), spec)
    assert all(torch.equal(back[k], sd[k]) for k in spec)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import gradtts_oracle as O
    cfg = UNetConfig()
    spec = estimator_param_spec(cfg)
    sd0 = synthetic_state_dict(cfg) if rank == 0 else None
    sd = broadcast_state_dict(sd0, spec, torch.device("cpu"))
    ref_sd = synthetic_state_dict(cfg)
    ok_weights = all(torch.equal(sd[k], ref_sd[k]) for k in spec)
    z, mask, mu, _, _ = synthetic_inputs(B, 16, ragged=True)
    calls = []

    def compute(zs, ms, mus, n, spk):
        calls.append(zs.shape[0])
        return O.reverse_diffusion(sd, cfg, zs, ms, mus, n)

    y = sharded_sample(compute, z, mask, mu, 2)
    full = O.reverse_diffusion(ref_sd, cfg, z, mask, mu, 2)
    q.put((rank, ok_weights, calls[0] if calls else 0, torch.allclose(y, full, rtol=1e-4, atol=1e-4 * full.abs().max().item()), tuple(y.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 3, 1])      # 1: fewer utterances than ranks - the empty rank still joins the gather
def test_two_rank_sharded_sampling_equals_single_process(B):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_w, nloc, ok_y, shape in res:
        assert ok_w, "broadcast weights differ"
        lo, hi = shard_bounds(B, world, rank)
        assert nloc == hi - lo
        assert ok_y and shape == (B, 80, 16)
