"""Engines of the sibling recommenders that share the n_items-wide decoder output layer with the AAE.

``DecoderEngine``  DecodingRecommender (aaerec/aae.py:461-584): the reference's ``Decoder`` (aae.py:149-178) alone, fed
                   with the concatenated condition encodings; BCE against the item sets; one Adam.
``VAEEngine``      VAE (aaerec/vae.py:47-266): fc1 (sparse first layer) -> relu -> fc21 | fc22 -> reparametrisation ->
                   [z | cond] -> fc3 -> relu -> fc4; BCE + KLD; one Adam over all parameters.

Both reuse ``AAEEngine``'s state, batch feed, CUDA-graph handling, the fused output-layer training kernel (K3,
``aae_dec_out_train_ws``) and the whole predict / ranking tail (K5); only the small-layer kernels differ
(``csrc/siblings.cu``).  Single GPU (the reference's siblings have no sharded counterpart to match).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from ._native import call, ptr, AaeDims
from .engine import AAEEngine, _on_device


def _t(params, name):
    v = params[name]
    return torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.float32)


class _SiblingEngine(AAEEngine):
    def __init__(self, *a, **kw):
        if int(kw.get("world", 1)) != 1:
            raise NotImplementedError("the sibling recommenders run on one GPU")
        kw["adversarial"] = False
        super().__init__(*a, **kw)

    def _zero_moments(self):
        for m in (self.W1_m1, self.W1_v1, self.W1_m2, self.W1_v2, self.Wd3_m, self.Wd3_v, self.bd3_m, self.bd3_v,
                  self.enc_m1, self.enc_v1, self.enc_m2, self.enc_v2, self.dec_m, self.dec_v, self.disc_m, self.disc_v):
            m.zero_()
        self._init_state()
        self.steps_done = 0
        self._wp_dirty = True

    def _load_block(self, blk, sizes, params):
        off = 0
        for name, sz in sizes:
            v = _t(params, name).reshape(-1)
            assert v.numel() == sz, (name, v.numel(), sz)
            blk[off:off + sz].copy_(v)
            off += sz

    def _export_block(self, blk, sizes, shapes, out):
        off = 0
        for name, sz in sizes:
            v = blk[off:off + sz].cpu()
            out[name] = v.reshape(shapes[name]) if name in shapes else v.clone()
            off += sz

    def init_uniform(self, seed=42):
        raise NotImplementedError("on-device random init exists for the AAE engine only; load_params a state dict")

    def make_replica(self, max_batch=1024):
        raise NotImplementedError("set-sharded predict replicas exist for the AAE engine only")

    def optim_state(self, which):
        raise NotImplementedError("optimizer-state export exists for the AAE engine only")


class DecoderEngine(_SiblingEngine):
    """``Decoder(code_size=D, n_hidden, n_items)`` of aae.py:149-178 as the whole model (aae.py:518-523): parameters
    ``lin1.*``, ``lin2.*`` (packed dec block) and ``lin3.*`` (Wd3 / bd3)."""
    has_encoder = False

    def __init__(self, n_items, n_hidden=100, cond_dim=0, lr=1e-3, dropout=(.2, .2), **kw):
        if cond_dim <= 0:
            raise AssertionError("Minimum 1 condition is necessary for MLP")        # aae.py:475
        super().__init__(n_items, n_hidden, 0, cond_dim=cond_dim, gen_lr=lr, reg_lr=0.0, dropout=dropout, **kw)

    def dec_sizes(self):
        H, D = self.H, self.D
        return [("lin1.weight", H * D), ("lin1.bias", H), ("lin2.weight", H * H), ("lin2.bias", H)]

    @_on_device
    def load_params(self, params):
        self.Wd3.copy_(_t(params, "lin3.weight"))
        self.bd3.copy_(_t(params, "lin3.bias"))
        self._load_block(self.dec, self.dec_sizes(), params)
        self._zero_moments()

    @_on_device
    def state_dict(self):
        torch.cuda.synchronize(self.dev)
        out = {"lin3.weight": self.Wd3[: self.Vloc].cpu().clone(), "lin3.bias": self.bd3[: self.Vloc].cpu().clone()}
        self._export_block(self.dec, self.dec_sizes(), {"lin1.weight": (self.H, self.D), "lin2.weight": (self.H, self.H)},
                           out)
        return out

    def flush_w1(self):
        self._w1_dirty = False          # no sparse first layer

    def _enqueue_ae(self, ctx):
        """DecodingRecommender.partial_fit (aae.py:490-520): mlp forward, BCE, backward, mlp_optim."""
        s = self._stream
        B, dims, dr = ctx["B"], ctx["dims"], ctx["dr"]
        st = ptr(self.state)
        call("aae_decoder_fwd", dims, ptr(self.cond), ptr(self.dec), dr["ae_d1"], dr["ae_d2"], st, ptr(self.dd1),
             ptr(self.h2), ptr(self.dh2), 1, s())
        self._dec_out_train(B, ctx["n_total"])
        call("aae_decoder_bwd", dims, ptr(self.dh2), ptr(self.dec), dr["ae_d1"], dr["ae_d2"], st, ptr(self.dd1),
             ptr(self.h2), ptr(self.g_d2), ptr(self.g_d1), s())
        call("aae_decoder_wgrad", dims, ptr(self.cond), ptr(self.dd1), ptr(self.g_d2), ptr(self.g_d1), None,
             N.adam_block(self.dec, self.dec_m, self.dec_v, 0), st, s())

    def _enqueue_finish(self, ctx):
        call("aae_step_finish", ptr(self.slot_of), ptr(self.uniq), ptr(self.n_uniq), ctx["cap"], ptr(self.loss_sums), 3,
             ctx["n_total"], ctx["B"], ptr(self.losses), ptr(self.state), ptr(self.ktab), self._stream())

    @_on_device
    def predict_h2(self, B):
        dims = AaeDims(B, self.H, 0, self.D)
        none = N.drop(None, 0.0, 0)
        call("aae_decoder_fwd", dims, ptr(self.cond), ptr(self.dec), none, none, None, None, ptr(self.h2), None, 0,
             self._stream())
        return self.h2[:B]


class VAEEngine(_SiblingEngine):
    """vae.py:47-101: parameters ``fc1`` (W1t / enc block), ``fc21`` / ``fc22`` (stacked in the enc block), ``fc3`` (dec
    block), ``fc4`` (Wd3 / bd3); one Adam at ``lr`` over all of them (vae.py:90-91)."""

    def __init__(self, n_items, n_hidden=100, n_code=50, cond_dim=0, lr=1e-3, **kw):
        self._vae_B = 0
        super().__init__(n_items, n_hidden, n_code, cond_dim=cond_dim, gen_lr=lr, reg_lr=0.0, dropout=(0.0, 0.0), **kw)

    def enc_sizes(self):
        H, Cc = self.H, self.C
        return [("fc1.bias", H), ("fcml.weight", 2 * Cc * H), ("fcml.bias", 2 * Cc)]

    def dec_sizes(self):
        return [("fc3.weight", self.H * self.Cp), ("fc3.bias", self.H)]

    def _ensure_ws(self, B, nnz):
        super()._ensure_ws(B, nnz)
        if self._vae_B < self._ws_B:
            f32 = dict(dtype=torch.float32, device=self.dev)
            Bw, Cc = self._ws_B, self.C
            self.mulv = torch.zeros(Bw, 2 * Cc, **f32)
            self.g_ml = torch.zeros(Bw, 2 * Cc, **f32)
            self.eps = torch.zeros(Bw, Cc, **f32)
            self.eps_used = torch.zeros(Bw, Cc, **f32)
            self._vae_B = Bw

    @_on_device
    def load_params(self, params):
        W1 = _t(params, "fc1.weight")                 # [H, V]
        assert W1.shape == (self.H, self.V), (W1.shape, self.H, self.V)
        self.W1t.copy_(W1.t().contiguous())
        self.Wd3.copy_(_t(params, "fc4.weight"))
        self.bd3.copy_(_t(params, "fc4.bias"))
        p = dict(params)
        p["fcml.weight"] = torch.cat([_t(params, "fc21.weight"), _t(params, "fc22.weight")], dim=0)
        p["fcml.bias"] = torch.cat([_t(params, "fc21.bias"), _t(params, "fc22.bias")], dim=0)
        self._load_block(self.enc, self.enc_sizes(), p)
        self._load_block(self.dec, self.dec_sizes(), p)
        self._zero_moments()

    @_on_device
    def state_dict(self):
        self.flush_w1()
        torch.cuda.synchronize(self.dev)
        H, Cc = self.H, self.C
        out = {"fc1.weight": self.W1t[: self.Vloc].t().contiguous().cpu(),
               "fc4.weight": self.Wd3[: self.Vloc].cpu().clone(), "fc4.bias": self.bd3[: self.Vloc].cpu().clone()}
        tmp = {}
        self._export_block(self.enc, self.enc_sizes(), {"fcml.weight": (2 * Cc, H)}, tmp)
        self._export_block(self.dec, self.dec_sizes(), {"fc3.weight": (H, self.Cp)}, tmp)
        out["fc1.bias"] = tmp["fc1.bias"]
        out["fc21.weight"], out["fc22.weight"] = tmp["fcml.weight"][:Cc].clone(), tmp["fcml.weight"][Cc:].clone()
        out["fc21.bias"], out["fc22.bias"] = tmp["fcml.bias"][:Cc].clone(), tmp["fcml.bias"][Cc:].clone()
        out["fc3.weight"], out["fc3.bias"] = tmp["fc3.weight"], tmp["fc3.bias"]
        return out

    last_B = 1      # rows of the last training step (losses[1] holds KLD / last_B)

    def train_step(self, B, injected=False):
        self.last_B = B
        return super().train_step(B, injected)

    def train_step_host(self, indptr_np, *a, **kw):
        self.last_B = int(indptr_np.shape[0]) - 1
        return super().train_step_host(indptr_np, *a, **kw)

    def phase_step(self, phase, B, injected=False):
        self.last_B = B
        return super().phase_step(phase, B, injected)

    @_on_device
    def set_eps(self, B, eps):
        """Oracle-RNG mode: the reference's ``torch.randn_like(std)`` draw of this batch (vae.py:109)."""
        self._ensure_ws(B, 0)
        self.eps[:B].copy_(torch.as_tensor(eps, dtype=torch.float32))

    def _enqueue_ae(self, ctx):
        """VAE.partial_fit (vae.py:147-186): forward with the reparametrisation, BCE + KLD, backward, Adam."""
        s = self._stream
        B, dims, bag, cap = ctx["B"], ctx["dims"], ctx["bag"], ctx["cap"]
        H, st = self.H, ptr(self.state)
        lo, hi = self.v_begin, self.v_end
        cur = torch.cuda.current_stream(self.dev)
        self._ev_fork.record(cur)
        self.side.wait_event(self._ev_fork)
        with torch.cuda.stream(self.side):
            call("aae_batch_prepare", ptr(self.indptr), ptr(self.indices), B, lo, hi, ptr(self.slot_of), ptr(self.uniq),
                 ptr(self.n_uniq), ptr(self.csc_cnt), ptr(self.csc_pos), ptr(self.csc_off), ptr(self.csc_row), cap, s())
            self._ev_prep.record(self.side)
        call("aae_w1_catchup", ptr(self.indptr), ptr(self.indices), B, lo, hi, ptr(self.w1_claim), ptr(self.W1t),
             ptr(self.W1_m1), ptr(self.W1_v1), ptr(self.W1_m2), ptr(self.W1_v2), ptr(self.w1_last), H, st,
             ptr(self.ktab), s())
        call("aae_vae_fwd", dims, bag, ptr(self.h1pre), ptr(self.cond), ptr(self.eps) if ctx["injected"] else None,
             ptr(self.enc), ptr(self.dec), st, ptr(self.a1), ptr(self.mulv), ptr(self.eps_used), ptr(self.zc),
             ptr(self.h2), ptr(self.dh2), ptr(self.loss_sums[1:]), s())
        self._dec_out_train(B, ctx["n_total"])
        if self.overlap_sweep:
            self._ev_k3.record(cur)
            self.side.wait_event(self._ev_k3)
            with torch.cuda.stream(self.side):
                self._sweep()
        call("aae_vae_bwd", dims, ptr(self.dh2), ptr(self.enc), ptr(self.dec), ptr(self.eps_used), ptr(self.a1),
             ptr(self.mulv), ptr(self.h2), ptr(self.g_d2), ptr(self.g_ml), ptr(self.g_h1), s())
        with self._branch():
            call("aae_vae_wgrad", dims, ptr(self.a1), ptr(self.zc), ptr(self.g_d2), ptr(self.g_ml), ptr(self.g_h1),
                 None, None, N.adam_block(self.enc, self.enc_m1, self.enc_v1, 0),
                 N.adam_block(self.dec, self.dec_m, self.dec_v, 0), st, s())
        cur.wait_event(self._ev_prep)
        call("aae_w1_rows_update", ptr(self.uniq), ptr(self.n_uniq), cap, ptr(self.csc_off), ptr(self.csc_row),
             ptr(self.indptr), self.normalize, ptr(self.g_h1), ptr(self.W1t), ptr(self.W1_m1), ptr(self.W1_v1), H, st,
             0, ptr(self.w1_last), s())
        self._join()

    eval_injected = False      # True: predict uses the noise set by set_eps (oracle-RNG mode), else in-kernel Philox

    @_on_device
    def predict_h2(self, B):
        """The reference's predict runs the full forward -- reparametrisation included, also in eval mode
        (vae.py:252-256 -> forward 116-124)."""
        self.flush_w1()
        dims = AaeDims(B, self.H, self.C, self.D)
        bag = N.bag(self.indptr, self.indices, self.W1t, self.normalize, self.v_begin, self.v_end)
        call("aae_vae_fwd", dims, bag, ptr(self.h1pre), ptr(self.cond), ptr(self.eps) if self.eval_injected else None,
             ptr(self.enc), ptr(self.dec), ptr(self.state), None, None, None, None, ptr(self.h2), None, None,
             self._stream())
        return self.h2[:B]
