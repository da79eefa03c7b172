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


class PeerExchange(object):
    """NVLink peer-memory exchange buffers of the ranks of one box (``csrc/peer.cu``).

    Every rank allocates one buffer through the C ABI (``aae_peer_alloc``: cudaMalloc + CUDA IPC handle), the
    handles are all-gathered over the process group, and each rank maps its peers' buffers
    (``aae_peer_open``).  ``allreduce(t, exchange)`` then enqueues ONE kernel that publishes ``t``, waits for the
    peers through flags in peer memory and sums the peers' copies in rank order -- no NCCL call, no host
    synchronisation, capturable in the step's CUDA graph.  Construction raises if IPC mapping is not possible
    (the engine then keeps NCCL all-reduces)."""

    def __init__(self, rank, world, group, n_max, device):
        import ctypes as C
        import torch.distributed as dist
        from . import _native as N
        if world > 8:
            raise RuntimeError("peer exchange supports up to 8 ranks of one box")
        self.rank, self.world, self.n_max = int(rank), int(world), int(n_max)
        self.device = device
        self._N = N
        lib = N.load()
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._own = None
        self._opened = []
        ok, msg = 1, ""
        try:
            N.call("aae_peer_alloc", self.n_max, C.byref(base), handle)
            self._own = base.value
        except Exception as e:   # noqa: BLE001 -- reported collectively below
            ok, msg = 0, repr(e)
        info = [None] * world
        dist.all_gather_object(info, (ok, bytes(handle.raw), msg), group=group)
        if not all(i[0] for i in info):
            self.close()
            raise RuntimeError("peer exchange: allocation failed on some rank: %r" % [i[2] for i in info])
        bases = [None] * world
        ok, msg = 1, ""
        for r in range(world):
            if r == rank:
                bases[r] = self._own
                continue
            p = C.c_void_p()
            try:
                N.call("aae_peer_open", info[r][1], C.byref(p))
                bases[r] = p.value
                self._opened.append(p.value)
            except Exception as e:   # noqa: BLE001
                ok, msg = 0, repr(e)
                break
        res = [None] * world
        dist.all_gather_object(res, (ok, msg), group=group)
        if not all(i[0] for i in res):
            self.close()
            raise RuntimeError("peer exchange: IPC mapping failed on some rank: %r" % [i[1] for i in res])
        arr = (C.c_void_p * 8)(*[C.c_void_p(b) for b in bases] + [None] * (8 - world))
        self.peers = N.AaePeers(arr, self.rank, self.world)
        self._lib = lib

    def allreduce(self, t, exchange, extra=None, stream=None):
        """In-place sum over the ranks of the contiguous float32 tensor ``t`` (and of up to 4 doubles ``extra``)."""
        import ctypes as C
        import torch
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.n_max
        if stream is None:
            stream = torch.cuda.current_stream(t.device).cuda_stream
        n_extra = 0 if extra is None else extra.numel()
        if extra is not None:
            assert extra.dtype == torch.float64 and extra.is_contiguous() and n_extra <= 4
        self._N.call("aae_peer_allreduce", self.peers, int(exchange), C.c_void_p(t.data_ptr()), t.numel(),
                     C.c_void_p(extra.data_ptr()) if extra is not None else None, n_extra, self.n_max,
                     C.c_void_p(stream))

    def error(self):
        """1 if a wait inside an exchange timed out (a peer died or the ranks diverged); synchronises."""
        import ctypes as C
        v = C.c_int(0)
        self._N.call("aae_peer_error", C.c_void_p(self._own), C.byref(v))
        return int(v.value)

    def close(self):
        for p in self._opened:
            self._N.call("aae_peer_close", C_void(p))
        self._opened = []
        if self._own:
            self._N.call("aae_peer_free", C_void(self._own))
            self._own = None


def C_void(p):
    import ctypes as C
    return C.c_void_p(p)
