"""Data-parallel sampling across the GPUs of one box (BASELINE config 5).

Utterances are independent (convs, GroupNorm, attention and the Euler update never mix batch
entries), so a batch shards into contiguous slices, one per rank, with NO collective inside the
N-step loop.  The only communication is what north_star names: a broadcast of the weights at
start-up and a gather of the output mels at the end.  One process per GPU, torch.distributed
(NCCL on GPUs; gloo in the CPU unit tests, which inject the compute function).
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous near-equal partition: the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_state_dict(sd: dict, names: list[str]) -> torch.Tensor:
    return torch.cat([sd[n].detach().reshape(-1).to(torch.float32) for n in names])


def unflatten_state_dict(flat: torch.Tensor, names: list[str], shapes: dict) -> dict:
    out, off = {}, 0
    for n in names:
        k = int(torch.Size(shapes[n]).numel())
        out[n] = flat[off:off + k].view(*shapes[n]).clone()
        off += k
    assert off == flat.numel()
    return out


def broadcast_state_dict(sd: dict | None, shapes: dict, device, src: int = 0) -> dict:
    """One flat broadcast of every estimator tensor (30.5 MB fp32 for Grad-TTS) from `src`."""
    names = list(shapes.keys())
    total = sum(int(torch.Size(s).numel()) for s in shapes.values())
    if dist.get_rank() == src:
        flat = flatten_state_dict(sd, names).to(device)
    else:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src)
    return unflatten_state_dict(flat, names, shapes)


def sharded_sample(compute: Callable, z, mask, mu, n_timesteps: int, spk=None, gather: bool = True):
    """Run `compute(z, mask, mu, n_timesteps, spk)` on this rank's slice of the batch and all-gather the mels.

    Every rank passes the FULL batch tensors (or at least its own slice's worth; only the slice is read).
    The padded length T is global: padding is semantically live (GroupNorm/attention reduce over padded
    columns), so all shards must use the batch's T, exactly as the single-GPU run does.
    Returns the full [B, n_feats, T] output on every rank when `gather`, else the local slice.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    B = z.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    sl = slice(lo, hi)
    if hi > lo:
        local = compute(z[sl], mask[sl], mu[sl], n_timesteps, None if spk is None else spk[sl])
    else:
        # fewer utterances than ranks: this rank has nothing to sample but must still take part in the gather
        local = z[sl].clone()
    if not gather:
        return local
    if B % world == 0:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged shards: pad to the largest shard, gather, then trim
    per = (B + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([parts[r][: shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0]] for r in range(world)])
