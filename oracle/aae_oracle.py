"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference AAE hot path.

This file is the *oracle*: a dense, autograd-free, float32 restatement of what
``/root/reference/aaerec/aae.py`` computes for one ``partial_fit`` (reconstruction,
discriminator and generator phases), for ``predict`` and for the ranking tail
(``remove_non_missing`` + ``argtopk``), plus the sibling models that share the decoder
output layer: AutoEncoder, DenoisingAutoEncoder (dae.py), DecodingRecommender (aae.py:461-584)
and VAE (vae.py:47-266).  It exists to check the CUDA kernels; it is
never imported by the product package.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import it.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(imported through ``oracle/reference_loader.py``) on seeded inputs and stores its
losses / weights / predictions under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those vectors and against the reference's own
doctest vectors for ``argtopk`` (evaluation.py:24-44) and ``remove_non_missing``
(evaluation.py:187-191).

The arithmetic lives in torch/ATen (un-vendored, unpinned by the reference's
setup.py:3-14; torch 2.11.0 in this image).  The published formulas restated here:
  * nn.Linear                     -> addmm(b, x, W^T)
  * F.normalize(x, p=1, dim=1)    -> x / max(sum|x|, 1e-12)          (aae.py:132-133)
  * nn.Dropout (train)            -> x * (bernoulli(1-p) / (1-p)), applied BEFORE ReLU
  * F.binary_cross_entropy        -> (t-1)*max(log1p(-x),-100) - t*max(log x,-100), mean;
                                     backward (x-t)/max((1-x)x, 1e-12)/N
  * optim.Adam (single tensor)    -> lerp_/mul_/addcmul_/sqrt/div/add_/addcdiv_ sequence
Every function cites the reference lines it follows.
"""
import math

import numpy as np
import torch

TINY = 1e-12  # aae.py:28

ENC = ("enc.lin1", "enc.lin2", "enc.lin3")
DEC = ("dec.lin1", "dec.lin2", "dec.lin3")
DISC = ("disc.lin1", "disc.lin2", "disc.lin3")


def _names(layers):
    out = []
    for l in layers:
        out += [l + ".weight", l + ".bias"]
    return out


def init_params(n_items, n_hidden=100, n_code=50, code_size=None, seed=42, adversarial=True):
    """Initial weights exactly as the reference builds them (aae.py:27, 782-792):
    ``torch.manual_seed(42)`` then Encoder(lin1, lin2, lin3) -> Decoder -> Discriminator,
    each ``nn.Linear`` with its stock init, on the CPU generator.  (``lin3`` of the
    encoder is constructed after the dropout modules, which draw nothing.)"""
    if code_size is None:
        code_size = n_code
    if seed is not None:
        torch.manual_seed(seed)
    shapes = [
        ("enc.lin1", n_items, n_hidden), ("enc.lin2", n_hidden, n_hidden), ("enc.lin3", n_hidden, n_code),
        ("dec.lin1", code_size, n_hidden), ("dec.lin2", n_hidden, n_hidden), ("dec.lin3", n_hidden, n_items),
        ("disc.lin1", n_code, n_hidden), ("disc.lin2", n_hidden, n_hidden), ("disc.lin3", n_hidden, 1),
    ]
    if not adversarial:     # AutoEncoder.fit builds Encoder and Decoder only (aae.py:368-387)
        shapes = shapes[:6]
    params = {}
    for name, fin, fout in shapes:
        lin = torch.nn.Linear(fin, fout)
        params[name + ".weight"] = lin.weight.detach().clone()
        params[name + ".bias"] = lin.bias.detach().clone()
    return params


class Adam(object):
    """torch.optim.Adam defaults (betas (0.9,0.999), eps 1e-8, no weight decay), the
    single-tensor CPU code path (torch/optim/adam.py::_single_tensor_adam); the
    reference builds four of them (aae.py:798-804)."""

    def __init__(self, names, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.names = list(names)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = {}
        self.v = {}

    def step(self, params, grads):
        self.t += 1
        bc1 = 1 - self.beta1 ** self.t
        bc2 = 1 - self.beta2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        for n in self.names:
            p, g = params[n], grads[n]
            if n not in self.m:
                self.m[n] = torch.zeros_like(p)
                self.v[n] = torch.zeros_like(p)
            m, v = self.m[n], self.v[n]
            m.lerp_(g, 1 - self.beta1)
            v.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


def _linear(x, W, b):
    return torch.addmm(b, x, W.t())


def _drop_relu(pre, mask):
    """Dropout is applied before the activation (aae.py:135-141, 168-174, 199-205)."""
    if mask is not None:
        pre = pre * mask
    return torch.relu(pre), pre


def _relu_drop_bwd(d_out, dropped, mask):
    d = d_out * (dropped > 0).to(d_out.dtype)
    if mask is not None:
        d = d * mask
    return d


def draw_masks(shape, p, n):
    """n dropout masks as torch's CPU dropout draws them: ``empty_like(x).bernoulli_(1-p)
    .div_(1-p)`` on the global CPU generator; p == 0 draws nothing."""
    if p == 0:
        return [None] * n
    return [torch.empty(shape, dtype=torch.float32).bernoulli_(1 - p).div_(1 - p) for _ in range(n)]


def draw_step_rng(B, n_hidden, n_code, dropout=(.2, .2), prior_scale=None, adversarial=True):
    """All random draws of one partial_fit in the reference's order:
    ae: enc.drop1, enc.drop2, dec.drop1, dec.drop2 | disc: randn[B,C], disc.drop1/2 on
    z_real, disc.drop1/2 on z_fake | gen: enc.drop1, enc.drop2, disc.drop1, disc.drop2."""
    p1, p2 = dropout

    def pair():
        a = draw_masks((B, n_hidden), p1, 1)[0]
        b = draw_masks((B, n_hidden), p2, 1)[0]
        return a, b
    r = {}
    r["ae_enc"] = pair()
    r["ae_dec"] = pair()
    if not adversarial:     # AutoEncoder.ae_step (aae.py:267-306): the four masks of the reconstruction phase only
        return r
    z_real = torch.randn((B, n_code))  # aae.py:716, CPU generator
    if prior_scale is not None:
        z_real = z_real * prior_scale
    r["z_real"] = z_real
    r["disc_real"] = pair()
    r["disc_fake"] = pair()
    r["gen_enc"] = pair()
    r["gen_disc"] = pair()
    return r


NO_DROPOUT = {k: (None, None) for k in ("ae_enc", "ae_dec", "disc_real", "disc_fake", "gen_enc", "gen_disc")}


class OracleAAE(object):
    """Dense restatement of AdversarialAutoEncoder (aae.py:589-870) for prior='gauss',
    activation='ReLU', optimizer='adam', concatenation conditions given as float matrices."""

    def __init__(self, params, n_code=50, gen_lr=0.001, reg_lr=0.001, normalize_inputs=True,
                 faithful_cost=False):
        self.p = {k: v.clone().float() for k, v in params.items()}
        self.n_code = n_code
        self.normalize_inputs = normalize_inputs
        self.faithful_cost = faithful_cost
        # aae.py:800-804 -- two optimizers share the encoder's parameters
        self.enc_optim = Adam(_names(ENC), gen_lr)
        self.dec_optim = Adam(_names(DEC), gen_lr)
        self.gen_optim = Adam(_names(ENC), reg_lr)
        self.disc_optim = Adam(_names(DISC), reg_lr)

    # ---- forward pieces -------------------------------------------------
    def _enc_fwd(self, X, masks):
        """Encoder.forward, aae.py:130-146 (final activation linear for the gauss prior)."""
        p = self.p
        if self.normalize_inputs:
            denom = X.abs().sum(1, keepdim=True).clamp_min(1e-12)
            Xn = X / denom
        else:
            Xn = X
        a1, dr1 = _drop_relu(_linear(Xn, p["enc.lin1.weight"], p["enc.lin1.bias"]), masks[0])
        a2, dr2 = _drop_relu(_linear(a1, p["enc.lin2.weight"], p["enc.lin2.bias"]), masks[1])
        z = _linear(a2, p["enc.lin3.weight"], p["enc.lin3.bias"])
        return z, (Xn, a1, dr1, a2, dr2)

    def _enc_bwd(self, dz, cache, masks, grads):
        p = self.p
        Xn, a1, dr1, a2, dr2 = cache
        grads["enc.lin3.weight"] = dz.t() @ a2
        grads["enc.lin3.bias"] = dz.sum(0)
        d2 = _relu_drop_bwd(dz @ p["enc.lin3.weight"], dr2, masks[1])
        grads["enc.lin2.weight"] = d2.t() @ a1
        grads["enc.lin2.bias"] = d2.sum(0)
        d1 = _relu_drop_bwd(d2 @ p["enc.lin2.weight"], dr1, masks[0])
        grads["enc.lin1.weight"] = d1.t() @ Xn          # dense [H,V], only set columns non-zero
        grads["enc.lin1.bias"] = d1.sum(0)

    def _mlp3_fwd(self, prefix, x, masks):
        p = self.p
        h1, dr1 = _drop_relu(_linear(x, p[prefix + ".lin1.weight"], p[prefix + ".lin1.bias"]), masks[0])
        h2, dr2 = _drop_relu(_linear(h1, p[prefix + ".lin2.weight"], p[prefix + ".lin2.bias"]), masks[1])
        out = torch.sigmoid(_linear(h2, p[prefix + ".lin3.weight"], p[prefix + ".lin3.bias"]))
        return out, (x, h1, dr1, h2, dr2)

    def _mlp3_bwd(self, prefix, d_pre3, cache, masks, grads, accumulate=False):
        """Backward of lin1->drop->relu->lin2->drop->relu->lin3 given dL/d(lin3 pre-activation)."""
        p = self.p
        x, h1, dr1, h2, dr2 = cache

        def put(name, val):
            if accumulate and name in grads:
                grads[name] = grads[name] + val
            else:
                grads[name] = val
        put(prefix + ".lin3.weight", d_pre3.t() @ h2)
        put(prefix + ".lin3.bias", d_pre3.sum(0))
        d2 = _relu_drop_bwd(d_pre3 @ p[prefix + ".lin3.weight"], dr2, masks[1])
        put(prefix + ".lin2.weight", d2.t() @ h1)
        put(prefix + ".lin2.bias", d2.sum(0))
        d1 = _relu_drop_bwd(d2 @ p[prefix + ".lin2.weight"], dr1, masks[0])
        put(prefix + ".lin1.weight", d1.t() @ x)
        put(prefix + ".lin1.bias", d1.sum(0))
        return d1 @ p[prefix + ".lin1.weight"]

    # ---- the three phases -----------------------------------------------
    def ae_step(self, X, cond, rng):
        """aae.py:676-711."""
        z, enc_cache = self._enc_fwd(X, rng["ae_enc"])
        zc = z if cond is None else torch.cat([z] + list(cond), dim=1)   # condition.py:90-99, 312-316
        x, dec_cache = self._mlp3_fwd("dec", zc, rng["ae_dec"])
        xin = x + TINY
        tin = X + TINY
        N = X.numel()
        # ATen binary_cross_entropy forward/backward (mean reduction)
        loss_el = (tin - 1) * torch.log1p(-xin).clamp_min(-100) - tin * torch.log(xin).clamp_min(-100)
        loss = loss_el.mean()
        dx = (xin - tin) / ((1 - xin) * xin).clamp_min(1e-12) / N
        dpre = dx * (1 - x) * x                                          # sigmoid backward
        grads = {}
        dzc = self._mlp3_bwd("dec", dpre, dec_cache, rng["ae_dec"], grads)
        self._enc_bwd(dzc[:, : self.n_code].contiguous(), enc_cache, rng["ae_enc"], grads)
        self.enc_optim.step(self.p, grads)
        self.dec_optim.step(self.p, grads)
        return float(loss)

    def disc_step(self, X, rng):
        """aae.py:713-732: encoder in eval mode (no dropout), only disc_optim steps."""
        B = X.shape[0]
        z_real = rng["z_real"]
        z_fake, enc_cache = self._enc_fwd(X, (None, None))
        d_real, c_real = self._mlp3_fwd("disc", z_real, rng["disc_real"])
        d_fake, c_fake = self._mlp3_fwd("disc", z_fake, rng["disc_fake"])
        a = d_real + TINY
        b = 1 - d_fake + TINY
        loss = -torch.mean(torch.log(a) + torch.log(b))
        dd_real = (-1.0 / B) / a
        dd_fake = (1.0 / B) / b
        grads = {}
        self._mlp3_bwd("disc", dd_real * (1 - d_real) * d_real, c_real, rng["disc_real"], grads)
        dz_fake = self._mlp3_bwd("disc", dd_fake * (1 - d_fake) * d_fake, c_fake, rng["disc_fake"], grads,
                                 accumulate=True)
        if self.faithful_cost:
            # the reference also back-propagates into the (non-detached) encoder and throws the
            # result away (aae.py:722, 739); only computed when timing the CPU baseline
            self._enc_bwd(dz_fake, enc_cache, (None, None), {})
        self.disc_optim.step(self.p, grads)
        return float(loss)

    def gen_step(self, X, rng):
        """aae.py:734-743: encoder in train mode, gen_optim (second Adam state) steps."""
        B = X.shape[0]
        z, enc_cache = self._enc_fwd(X, rng["gen_enc"])
        d, c = self._mlp3_fwd("disc", z, rng["gen_disc"])
        a = d + TINY
        loss = -torch.mean(torch.log(a))
        dd = (-1.0 / B) / a
        dz = self._mlp3_bwd("disc", dd * (1 - d) * d, c, rng["gen_disc"], {})
        grads = {}
        self._enc_bwd(dz, enc_cache, rng["gen_enc"], grads)
        self.gen_optim.step(self.p, grads)
        return float(loss)

    def partial_fit(self, X, cond=None, rng=None):
        """aae.py:745-766.  X: dense [B,V] float32 (0/1); cond: list of [B,D_i] float32;
        rng: dict from draw_step_rng (or NO_DROPOUT + 'z_real')."""
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        if cond is not None:
            cond = [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond]
        r = self.ae_step(X, cond, rng)
        d = self.disc_step(X, rng)
        g = self.gen_step(X, rng)
        return r, d, g

    def predict(self, X, cond=None):
        """aae.py:840-870 for one batch: eval mode, sigmoid probabilities."""
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        z, _ = self._enc_fwd(X, (None, None))
        if cond is not None:
            z = torch.cat([z] + [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond], dim=1)
        x, _ = self._mlp3_fwd("dec", z, (None, None))
        return x.numpy()

    def logits(self, X, cond=None):
        """Pre-sigmoid decoder output (what the fused top-k ranks on)."""
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        z, _ = self._enc_fwd(X, (None, None))
        if cond is not None:
            z = torch.cat([z] + [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond], dim=1)
        _, cache = self._mlp3_fwd("dec", z, (None, None))
        h2 = cache[3]
        return _linear(h2, self.p["dec.lin3.weight"], self.p["dec.lin3.bias"]).numpy()


class OracleAE(OracleAAE):
    """Dense restatement of the plain AutoEncoder (aae.py:221-458): ``ae_step`` alone (aae.py:267-306 is the same
    computation as AdversarialAutoEncoder.ae_step), enc_optim and dec_optim both at ``lr`` (aae.py:393-394)."""

    def __init__(self, params, n_code=50, lr=0.001, normalize_inputs=True):
        params = dict(params)
        for name, shape in (("disc.lin1.weight", (1, n_code)), ("disc.lin1.bias", (1,)), ("disc.lin2.weight", (1, 1)),
                            ("disc.lin2.bias", (1,)), ("disc.lin3.weight", (1, 1)), ("disc.lin3.bias", (1,))):
            params.pop(name, None)
        OracleAAE.__init__(self, params, n_code=n_code, gen_lr=lr, reg_lr=lr, normalize_inputs=normalize_inputs)
        self.gen_optim = self.disc_optim = None

    def partial_fit(self, X, cond=None, rng=None):
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        if cond is not None:
            cond = [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond]
        return (self.ae_step(X, cond, rng),)


class OracleDAE(OracleAE):
    """Denoising autoencoder (dae.py:144-314): the plain autoencoder step on a batch whose entries were zeroed IN PLACE
    with probability noise_factor (dae.py:48-52, 191) -- input and BCE target are both the thinned batch
    (dae.py:198-200).  ``rng['noise']`` holds the reference's ``torch.rand(batch.size())`` draw."""

    def __init__(self, params, n_code=50, lr=0.001, normalize_inputs=True, noise_factor=0.2):
        super().__init__(params, n_code=n_code, lr=lr, normalize_inputs=normalize_inputs)
        self.noise_factor = noise_factor

    def partial_fit(self, X, cond=None, rng=None):
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32).clone()
        X[rng["noise"] < self.noise_factor] = 0
        return super().partial_fit(X, cond, rng)


def draw_dae_rng(B, V, n_hidden, n_code, dropout=(.2, .2)):
    """DAE draws of one step in the reference's order: torch.rand(batch.size()) (dae.py:50), then the four dropout masks
    of the reconstruction phase."""
    r = {"noise": torch.rand((B, V))}
    r.update(draw_step_rng(B, n_hidden, n_code, dropout, adversarial=False))
    return r


class OracleDecoder(object):
    """DecodingRecommender (aae.py:461-584): the reference's ``Decoder`` (aae.py:149-178: lin1 -> drop -> relu -> lin2 ->
    drop -> relu -> lin3 -> sigmoid) on the concatenated condition encodings (aae.py:495-507), BCE(y_pred + TINY,
    y + TINY) (aae.py:510), one Adam at ``lr`` over the mlp's parameters (aae.py:522-523).  ``params``: lin1/lin2/lin3."""

    def __init__(self, params, lr=0.001):
        self.net = OracleAAE.__new__(OracleAAE)
        self.net.p = {"dec." + k: torch.as_tensor(np.asarray(v)).clone().float() for k, v in params.items()}
        self.optim = Adam(_names(DEC), lr)

    @property
    def p(self):
        return {k[4:]: v for k, v in self.net.p.items()}

    def partial_fit(self, cond, Y, rng=None):
        """aae.py:490-520.  cond: list of [B, D_i] float matrices (concatenated in order); Y dense [B,V] 0/1;
        rng['ae_dec'] = the two dropout masks (Decoder.drop1, drop2)."""
        inp = torch.cat([torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond], dim=1)
        Y = torch.as_tensor(np.asarray(Y), dtype=torch.float32)
        masks = (rng or NO_DROPOUT)["ae_dec"]
        x, cache = self.net._mlp3_fwd("dec", inp, masks)
        xin, tin, N = x + TINY, Y + TINY, Y.numel()
        loss = ((tin - 1) * torch.log1p(-xin).clamp_min(-100) - tin * torch.log(xin).clamp_min(-100)).mean()
        dx = (xin - tin) / ((1 - xin) * xin).clamp_min(1e-12) / N
        grads = {}
        self.net._mlp3_bwd("dec", dx * (1 - x) * x, cache, masks, grads)
        self.optim.step(self.net.p, grads)
        return float(loss)

    def predict(self, cond):
        """aae.py:555-584 for one batch (eval mode)."""
        inp = torch.cat([torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond], dim=1)
        return self.net._mlp3_fwd("dec", inp, (None, None))[0].numpy()


VAE_LAYERS = ("fc1", "fc21", "fc22", "fc3", "fc4")


class OracleVAE(object):
    """VAE (vae.py:47-266), dense and autograd-free: h1 = relu(fc1(normalize(x))); mu, logvar = fc21(h1), fc22(h1);
    z = eps * exp(logvar / 2) + mu (vae.py:103-110); conditions concatenated on z (vae.py:120-123); recon =
    sigmoid(fc4(relu(fc3(z)))); loss = BCELoss()(recon, x) [mean: the later ``size_average = False`` assignment is
    inert, vae.py:127-130] + KLD, KLD = -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) (vae.py:136-143); one Adam over all
    parameters (vae.py:90-91).  ``eps`` = the step's ``torch.randn_like(std)`` draw."""

    def __init__(self, params, n_code=50, lr=0.001, normalize_inputs=True):
        self.p = {k: torch.as_tensor(np.asarray(v)).clone().float() for k, v in params.items()}
        self.n_code = n_code
        self.normalize_inputs = normalize_inputs
        self.optim = Adam(_names(VAE_LAYERS), lr)

    def _fwd(self, X, cond, eps):
        p = self.p
        Xn = X / X.abs().sum(1, keepdim=True).clamp_min(1e-12) if self.normalize_inputs else X
        pre1 = _linear(Xn, p["fc1.weight"], p["fc1.bias"])
        h1 = torch.relu(pre1)
        mu = _linear(h1, p["fc21.weight"], p["fc21.bias"])
        lv = _linear(h1, p["fc22.weight"], p["fc22.bias"])
        std = (lv * 0.5).exp()
        z = eps * std + mu
        zc = z if cond is None else torch.cat([z] + list(cond), dim=1)
        pre3 = _linear(zc, p["fc3.weight"], p["fc3.bias"])
        h3 = torch.relu(pre3)
        x = torch.sigmoid(_linear(h3, p["fc4.weight"], p["fc4.bias"]))
        return x, (Xn, pre1, h1, mu, lv, std, zc, pre3, h3)

    def partial_fit(self, X, cond=None, eps=None):
        """vae.py:147-186; returns the step's loss (BCE mean + KLD sum) as the reference computes it."""
        p = self.p
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        eps = torch.as_tensor(np.asarray(eps), dtype=torch.float32)
        if cond is not None:
            cond = [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond]
        x, (Xn, pre1, h1, mu, lv, std, zc, pre3, h3) = self._fwd(X, cond, eps)
        N = X.numel()
        bce = ((X - 1) * torch.log1p(-x).clamp_min(-100) - X * torch.log(x).clamp_min(-100)).mean()
        kld = -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
        dpre4 = (x - X) / ((1 - x) * x).clamp_min(1e-12) / N * (1 - x) * x
        g = {"fc4.weight": dpre4.t() @ h3, "fc4.bias": dpre4.sum(0)}
        dpre3 = (dpre4 @ p["fc4.weight"]) * (pre3 > 0).float()
        g["fc3.weight"], g["fc3.bias"] = dpre3.t() @ zc, dpre3.sum(0)
        dz = (dpre3 @ p["fc3.weight"])[:, : self.n_code]
        dmu = dz + mu
        dlv = dz * eps * 0.5 * std + 0.5 * (lv.exp() - 1)
        g["fc21.weight"], g["fc21.bias"] = dmu.t() @ h1, dmu.sum(0)
        g["fc22.weight"], g["fc22.bias"] = dlv.t() @ h1, dlv.sum(0)
        dpre1 = (dmu @ p["fc21.weight"] + dlv @ p["fc22.weight"]) * (pre1 > 0).float()
        g["fc1.weight"], g["fc1.bias"] = dpre1.t() @ Xn, dpre1.sum(0)
        self.optim.step(p, g)
        return float(bce + kld)

    def predict(self, X, cond=None, eps=None):
        """vae.py:231-266 for one batch: the full forward, sampling included."""
        X = torch.as_tensor(np.asarray(X), dtype=torch.float32)
        eps = torch.as_tensor(np.asarray(eps), dtype=torch.float32)
        if cond is not None:
            cond = [torch.as_tensor(np.asarray(c), dtype=torch.float32) for c in cond]
        return self._fwd(X, cond, eps)[0].numpy()


def fit_epoch_order(n):
    """Row order of one epoch: ``sklearn.utils.shuffle(X)`` (aae.py:815-817) with
    random_state=None permutes ``arange(n)`` with the global numpy generator."""
    idx = np.arange(n)
    np.random.shuffle(idx)
    return idx


# ---- ranking tail -----------------------------------------------------------
def minmax_scale_rows(Y):
    """sklearn.preprocessing.minmax_scale(Y, (0,1), axis=1) as used at evaluation.py:193:
    per row scale = 1/(max-min) (1 when the range is 0 / below 10*eps), X*scale + (0-min*scale),
    in the input's float dtype."""
    Y = np.array(Y, dtype=Y.dtype if np.issubdtype(np.asarray(Y).dtype, np.floating) else np.float64)
    mn = Y.min(axis=1)
    mx = Y.max(axis=1)
    rng = mx - mn
    rng = np.where(rng < 10 * np.finfo(rng.dtype).eps, 1.0, rng).astype(Y.dtype)
    scale = (1.0 / rng).astype(Y.dtype)
    mn_ = (0.0 - mn * scale).astype(Y.dtype)
    out = Y * scale[:, None]
    out += mn_[:, None]
    return out


def remove_non_missing(Y_pred, X_test):
    """evaluation.py:183-199."""
    Ys = minmax_scale_rows(np.asarray(Y_pred))
    Ys[X_test.nonzero()] = 0.
    return Ys


def argtopk(X, k):
    """evaluation.py:20-58 (note the ``k >= X.size`` test on the whole array, line 48)."""
    X = np.asarray(X)
    assert X.ndim == 2
    rows = np.arange(X.shape[0])[:, np.newaxis]
    if k is None or k >= X.size:
        return rows, np.argsort(X, axis=1)[:, ::-1]
    assert k > 0
    ind = np.argpartition(X, -k, axis=1)[:, -k:]
    cols = ind[rows, np.argsort(X[rows, ind], axis=1)][:, ::-1]
    return rows, cols


def rank_topk(Y_pred, X_test, k):
    """The consumer chain evaluation.py:375 -> 388 / make_submission.py:36-53."""
    return argtopk(remove_non_missing(Y_pred, X_test), k)[1]


# ---- ranking metrics (evaluation.py:70-164, 202-240; rank_metrics_with_std.py:13-40, 108-154) ----------------
def relevance_in_rank_order(y_true, y_pred, k):
    """RankingMetric.__call__ (evaluation.py:81-92): gold flags looked up in the top-k order of the prediction."""
    ind = argtopk(np.asarray(y_pred), k)
    return np.asarray(y_true)[ind]


def _reciprocal_rank(r):
    z = np.asarray(r).nonzero()[0]
    return 1.0 / (z[0] + 1) if z.size else 0.0


def _average_precision(r):
    r = np.asarray(r) != 0
    out = [np.mean(r[:i + 1]) for i in range(r.size) if r[i]]
    return float(np.mean(out)) if out else 0.0


def metric_per_row(name, y_true, y_pred):
    """One metric of evaluation.py:166-180 ('mrr@5', 'map', 'P@1' ...) per row of a dense (gold, prediction) pair."""
    name = name.lower()
    kind, _, kk = name.partition("@")
    k = int(kk) if kk else None
    rs = relevance_in_rank_order(y_true, y_pred, k)
    if kind == "mrr":
        return np.array([_reciprocal_rank(r) for r in rs])
    if kind == "map":
        return np.array([_average_precision(r) for r in rs])
    if kind == "p":
        return (rs > 0).mean(axis=1)
    raise KeyError(name)


def evaluate(y_true, y_pred, metrics):
    """evaluation.py:202-240 (batch_size=None): [(mean, std)] per metric."""
    y_true = y_true.toarray() if hasattr(y_true, "toarray") else np.asarray(y_true)
    out = []
    for m in metrics:
        v = metric_per_row(m, y_true, y_pred)
        out.append((float(v.mean()), float(v.std())))
    return out
