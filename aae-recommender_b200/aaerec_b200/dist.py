"""Item-sharding helpers shared by the engine and the (CPU, gloo) tests.

The decoder output layer ``Wd3/bd3`` and the encoder's first layer ``W1t`` are split by contiguous
item ranges over the ranks of one box; the small layers are replicated and computed
redundantly.  Per train step the only exchanges are all-reduce(sum) of the [B,H] partial sums of
``X.W1^T`` (twice) and of ``dh2`` (+ the loss partial), per predict batch an all-gather of the
per-shard top-k candidates.
"""
import torch


def shard_range(V, rank, world):
    """Contiguous item range of ``rank``: ceil-divided so every rank but the last is equal."""
    per = (V + world - 1) // world
    lo = min(V, rank * per)
    return lo, min(V, lo + per)


def gather_item_shards(local, V, world, group=None):
    """All-gather an item-sharded [Vloc, ...] tensor into the full [V, ...] tensor (state export and
    the API-compatible dense ``predict``).  Works for CPU (gloo) and CUDA (nccl) tensors."""
    if world == 1:
        return local
    import torch.distributed as dist
    per = (V + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat(out, 0)[:V]


def gather_topk_candidates(val, idx, k_pad, world, group=None):
    """All-gather per-shard top-k (val [B,kl] descending, idx [B,kl] global ids) padded to ``k_pad``
    columns -> candidate lists [B, world*k_pad] for the k-way merge."""
    import torch.distributed as dist
    B, kl = val.shape
    pv = torch.full((B, k_pad), -3.0e38, dtype=val.dtype, device=val.device)
    pi = torch.full((B, k_pad), -1, dtype=idx.dtype, device=idx.device)
    pv[:, :kl] = val
    pi[:, :kl] = idx
    gv = [torch.empty_like(pv) for _ in range(world)]
    gi = [torch.empty_like(pi) for _ in range(world)]
    dist.all_gather(gv, pv, group=group)
    dist.all_gather(gi, pi, group=group)
    return torch.cat(gv, 1).contiguous(), torch.cat(gi, 1).contiguous()
