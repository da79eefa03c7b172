"""world_size-2 (and 3) gloo tests on CPU of the item-sharding logic: shard ranges, the exchanges the
engine issues (all-reduce of X.W1^T partial sums, of dh2 and of the loss partial; all-gather of shards and
of top-k candidates).  The per-shard arithmetic here is the CPU oracle's -- on the GPU box the same
exchanges wrap the CUDA kernels (engine.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "aae-recommender_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aaerec_b200.dist import shard_range, gather_item_shards, gather_topk_candidates
        from aaerec_b200.synth import synth_sets
        from oracle import aae_oracle as O
        V, H, C, B, k = 1003, 16, 8, 20, 7
        X = torch.as_tensor(synth_sets(B, V, 6, seed=3).toarray())
        p = O.init_params(V, H, C, seed=5)
        lo, hi = shard_range(V, rank, world)
        # (1) encoder first layer: partial sums over the local items, bias on rank 0 only
        Xn = X / X.abs().sum(1, keepdim=True).clamp_min(1e-12)
        part = Xn[:, lo:hi] @ p["enc.lin1.weight"][:, lo:hi].t()
        if rank == 0:
            part = part + p["enc.lin1.bias"]
        dist.all_reduce(part)
        full = torch.addmm(p["enc.lin1.bias"], Xn, p["enc.lin1.weight"].t())
        assert torch.allclose(part, full, atol=1e-6)
        # (2) decoder output layer: local logits -> local loss sum, local dZ, dh2 partial; one packed all-reduce
        h2 = torch.relu(torch.randn(B, H, generator=torch.Generator().manual_seed(1)))
        Wl, bl = p["dec.lin3.weight"][lo:hi], p["dec.lin3.bias"][lo:hi]
        x = torch.sigmoid(torch.addmm(bl, h2, Wl.t()))
        t = X[:, lo:hi]
        loss_part = torch.nn.functional.binary_cross_entropy(x + 1e-12, t + 1e-12, reduction="sum")
        dz = (x - t) / (B * V)
        packed = torch.cat([(dz @ Wl).reshape(-1), loss_part.reshape(1)])
        dist.all_reduce(packed)
        xf = torch.sigmoid(torch.addmm(p["dec.lin3.bias"], h2, p["dec.lin3.weight"].t()))
        loss_full = torch.nn.functional.binary_cross_entropy(xf + 1e-12, X + 1e-12, reduction="sum")
        dh2_full = ((xf - X) / (B * V)) @ p["dec.lin3.weight"]
        assert torch.allclose(packed[:-1].reshape(B, H), dh2_full, atol=1e-7)
        assert abs(packed[-1].item() - loss_full.item()) / loss_full.item() < 1e-5
        # (3) shard gather (state export): rows come back in item order
        g = gather_item_shards(p["dec.lin3.weight"][lo:hi].clone(), V, world)
        assert torch.equal(g, p["dec.lin3.weight"])
        # (4) predict: per-shard masked top-k with global ids, all-gather, merge
        scores = torch.addmm(p["dec.lin3.bias"], h2, p["dec.lin3.weight"].t())
        loc = scores[:, lo:hi].clone()
        loc[X[:, lo:hi] > 0] = -3.0e38
        kl = min(k, hi - lo)
        v, i = torch.topk(loc, kl, dim=1)
        cv, ci = gather_topk_candidates(v, (i + lo).to(torch.int32), min(k, (V + world - 1) // world), world)
        mv, mi = torch.topk(cv, k, dim=1)
        merged = torch.gather(ci, 1, mi).numpy().astype(np.int64)
        want = O.rank_topk(scores.numpy(), X.numpy(), k)
        assert np.array_equal(merged, want)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_item_sharding_exchanges_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
