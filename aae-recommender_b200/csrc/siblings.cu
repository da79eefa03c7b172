// Small-layer kernels of the sibling recommenders that share the n_items-wide decoder output layer (K3 / K5) with the
// AAE: the decoder-only DecodingRecommender (aaerec/aae.py:461-584, its `Decoder` 149-178 fed with the
// concatenated condition encodings) and the variational autoencoder (aaerec/vae.py:47-266).  Same construction as
// mlp.cu: a CTA owns R rows of the batch and walks the whole chain with the activations in shared memory, the
// (L2-resident) weights streamed through two staging buffers; weight gradients + Adam are jobs of small_wgrad_kernel.
#include "mlp_blocks.cuh"

namespace aae {

// ---------------------------------------------------------------------------------------------
// DecodingRecommender: inp [B,D] -> lin1 -> drop -> relu -> lin2 -> drop -> relu = h2 (aae.py:165-175); lin3 is K3.
// dec block: [Wd1 (H*D) | bd1 (H) | Wd2 (H*H) | bd2 (H)].  Shared memory: [stage 0 | stage 1 | x | y]
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) decoder_fwd_kernel(aae_dims d, const float* __restrict__ inp,
                                                                  const float* __restrict__ dec, aae_drop d1,
                                                                  aae_drop d2, const aae_step_state* st, float* dd1,
                                                                  float* h2, float* dh2_zero, int train) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[2];
  const int H = d.H, D = d.D, B = d.B;
  const int ld = max(H, D);
  float* x = sm + 2 * STAGE_FLOATS;
  float* y = x + R * ld;
  const int row0 = blockIdx.x * R;
  if (dh2_zero)
    for (int q = threadIdx.x; q < R * H; q += blockDim.x)
      if (row0 + q / H < B) dh2_zero[(size_t)row0 * H + q] = 0.f;
  DecBlock Dc(dec, H, D);
  aae_drop none = {nullptr, 0.f, 0};
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(Dc.Wd1, H, D);
    layers[1] = Stager::make_layer(Dc.Wd2, H, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 2);
  load_rows<R>(x, ld, inp, D, row0, B);
  __syncthreads();
  layer_fwd<R>(sg, 0, x, ld, Dc.bd1, y, ld);
  drop_relu<R>(y, ld, H, row0, B, train ? d1 : none, st, dd1);
  layer_fwd<R>(sg, 1, y, ld, Dc.bd2, x, ld);
  drop_relu<R>(x, ld, H, row0, B, train ? d2 : none, st, h2);
}

// backward from dh2 = dL/dh2 (written by K3) to the pre-activation gradients g_d2, g_d1.  Shared: [stages | g | t | act]
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) decoder_bwd_kernel(aae_dims d, const float* __restrict__ dh2,
                                                                  const float* __restrict__ dec, aae_drop d1,
                                                                  aae_drop d2, const aae_step_state* st,
                                                                  const float* __restrict__ dd1,
                                                                  const float* __restrict__ h2, float* g_d2,
                                                                  float* g_d1) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[1];
  const int H = d.H, D = d.D, B = d.B;
  const int ld = max(H, D);
  float* g = sm + 2 * STAGE_FLOATS;
  float* t = g + R * ld;
  float* act = t + R * ld;
  const int row0 = blockIdx.x * R;
  DecBlock Dc(dec, H, D);
  if (threadIdx.x == 0) layers[0] = Stager::make_layer(Dc.Wd2, H, H);
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 1);
  load_rows<R>(g, ld, dh2, H, row0, B);
  load_rows<R>(act, ld, h2, H, row0, B);
  drop_relu_bwd<R>(g, act, ld, H, row0, B, d2, st, g_d2);
  layer_bwd<R>(sg, 0, g, ld, t, ld);
  load_rows<R>(act, ld, dd1, H, row0, B);
  drop_relu_bwd<R>(t, act, ld, H, row0, B, d1, st, g_d1);
}

// ---------------------------------------------------------------------------------------------
// VAE (vae.py:103-124): h1 = relu(fc1(normalize(x))) ; (mu | logvar) = fc21/fc22(h1) ; z = eps * exp(logvar / 2) + mu ;
// zc = [z | cond] ; h3 = relu(fc3(zc)) ; fc4 + sigmoid + BCE is K3.  No dropout (vae.py:59-60).
// enc block: [b1 (H) | Wml (2C*H) | bml (2C)]  -- rows 0..C-1 of Wml are fc21 (mu), C..2C-1 fc22 (logvar)
// dec block: [W3 (H*Cp) | b3 (H)]
// KLD = -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) (vae.py:132-145) is accumulated into kld_sum[0].
// ---------------------------------------------------------------------------------------------
struct VaeEnc {
  const float *b1, *Wml, *bml;
  __device__ VaeEnc(const float* p, int H, int C) { b1 = p; Wml = b1 + H; bml = Wml + (size_t)2 * C * H; }
};
struct VaeDec {
  const float *W3, *b3;
  __device__ VaeDec(const float* p, int H, int Cp) { W3 = p; b3 = W3 + (size_t)H * Cp; }
};

template <int R>
__global__ void __launch_bounds__(MLP_THREADS) vae_fwd_kernel(aae_dims d, aae_bag bag, const float* __restrict__ h1pre,
                                                              const float* __restrict__ cond,
                                                              const float* __restrict__ eps,
                                                              const float* __restrict__ enc,
                                                              const float* __restrict__ dec, const aae_step_state* st,
                                                              float* a1, float* mulv, float* eps_used, float* zc,
                                                              float* h3, float* dh3_zero, double* kld_sum) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[2];
  __shared__ float red[MLP_THREADS / 32];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(max(H, Cp), 2 * C);
  float* x = sm + 2 * STAGE_FLOATS;
  float* y = x + R * ld;
  float* scratch = align16f(y + R * ld);
  const int row0 = blockIdx.x * R;
  if (dh3_zero)
    for (int q = threadIdx.x; q < R * H; q += blockDim.x)
      if (row0 + q / H < B) dh3_zero[(size_t)row0 * H + q] = 0.f;
  VaeEnc E(enc, H, C);
  VaeDec Dc(dec, H, Cp);
  aae_drop none = {nullptr, 0.f, 0};
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(E.Wml, 2 * C, H);
    layers[1] = Stager::make_layer(Dc.W3, H, Cp);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 2);
  input_rows<R>(x, ld, bag, h1pre, E.b1, H, row0, B, scratch);
  drop_relu<R>(x, ld, H, row0, B, none, st, a1);
  layer_fwd<R>(sg, 0, x, ld, E.bml, y, ld);          // y[0..C) = mu, y[C..2C) = logvar
  // reparametrize (vae.py:107-110) + KLD partial sum; x[0..C) <- z
  float kld = 0.f;
  for (int q = threadIdx.x; q < R * C; q += blockDim.x) {
    const int r = q / C, i = q - r * C, row = row0 + r;
    float z = 0.f;
    if (row < B) {
      const float mu = y[r * ld + i], lv = y[r * ld + C + i];
      const float e = eps ? eps[(size_t)row * C + i] : randn_elem(st, (uint32_t)(row * C + i), 91u);
      const float sd = expf(0.5f * lv);
      z = fmaf(e, sd, mu);
      kld += -0.5f * (1.0f + lv - mu * mu - expf(lv));
      if (eps_used) eps_used[(size_t)row * C + i] = e;
    }
    x[r * ld + i] = z;
  }
  if (mulv) store_rows<R>(y, ld, mulv, 2 * C, row0, B);
  for (int q = threadIdx.x; q < R * d.D; q += blockDim.x) {       // condition.py:312-316: concat on the code
    const int r = q / d.D, i = q - r * d.D, row = row0 + r;
    x[r * ld + C + i] = (row < B) ? cond[(size_t)row * d.D + i] : 0.f;
  }
  __syncthreads();
  if (zc) store_rows<R>(x, ld, zc, Cp, row0, B);
  layer_fwd<R>(sg, 1, x, ld, Dc.b3, y, ld);
  drop_relu<R>(y, ld, H, row0, B, none, st, h3);
  if (kld_sum) {
    kld = warp_sum(kld);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kld;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < MLP_THREADS / 32; ++w) s += red[w];
      atomicAdd(kld_sum, (double)s);
    }
  }
}

// backward from dh3 (K3) to g_3 (fc3 pre-activation), g_ml = (dmu | dlogvar) and g_h1 (fc1 pre-activation).
// dmu = dz + mu ; dlogvar = dz * eps * exp(logvar/2) / 2 + (exp(logvar) - 1) / 2   (autograd of vae.py:107-110, 142-143)
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) vae_bwd_kernel(aae_dims d, const float* __restrict__ dh3,
                                                              const float* __restrict__ enc,
                                                              const float* __restrict__ dec,
                                                              const float* __restrict__ eps_used,
                                                              const float* __restrict__ a1,
                                                              const float* __restrict__ mulv,
                                                              const float* __restrict__ h3, float* g_3, float* g_ml,
                                                              float* g_h1) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[2];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(max(H, Cp), 2 * C);
  float* g = sm + 2 * STAGE_FLOATS;
  float* t = g + R * ld;
  float* act = t + R * ld;
  const int row0 = blockIdx.x * R;
  VaeEnc E(enc, H, C);
  VaeDec Dc(dec, H, Cp);
  aae_drop none = {nullptr, 0.f, 0};
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(Dc.W3, H, Cp);
    layers[1] = Stager::make_layer(E.Wml, 2 * C, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 2);
  load_rows<R>(g, ld, dh3, H, row0, B);
  load_rows<R>(act, ld, h3, H, row0, B);
  drop_relu_bwd<R>(g, act, ld, H, row0, B, none, nullptr, g_3);
  layer_bwd<R>(sg, 0, g, ld, t, ld);                 // t[0..Cp) = d(zc); the first C entries are dz
  for (int q = threadIdx.x; q < R * C; q += blockDim.x) {
    const int r = q / C, i = q - r * C, row = row0 + r;
    float dmu = 0.f, dlv = 0.f;
    if (row < B) {
      const float dz = t[r * ld + i];
      const float mu = mulv[(size_t)row * 2 * C + i], lv = mulv[(size_t)row * 2 * C + C + i];
      const float e = eps_used[(size_t)row * C + i];
      dmu = dz + mu;
      dlv = fmaf(dz * e, 0.5f * expf(0.5f * lv), 0.5f * (expf(lv) - 1.0f));
    }
    g[r * ld + i] = dmu;
    g[r * ld + C + i] = dlv;
  }
  __syncthreads();
  store_rows<R>(g, ld, g_ml, 2 * C, row0, B);
  layer_bwd<R>(sg, 1, g, ld, t, ld);                 // t[0..H) = d(h1)
  load_rows<R>(act, ld, a1, H, row0, B);
  drop_relu_bwd<R>(t, act, ld, H, row0, B, none, nullptr, g_h1);
}

}  // namespace aae

using namespace aae;

extern "C" {

int aae_decoder_fwd(aae_dims d, const float* inp, const float* dec, aae_drop d1, aae_drop d2, const aae_step_state* st,
                    float* dd1, float* h2, float* dh2_zero, int train, void* stream) {
  AAE_REQUIRE(inp && dec && h2, "null pointer");
  AAE_REQUIRE(!train || (st && dd1), "training needs the step state and the dd1 buffer");
  AAE_REQUIRE(d.B > 0 && d.H > 0 && d.C == 0 && d.D > 0 && d.H <= 2048 && d.D <= 4096, "size outside envelope");
  int ld = std::max(d.H, d.D);
  LAUNCH_R(decoder_fwd_kernel, d.B, 2 * ld, stream, d, inp, dec, d1, d2, st, dd1, h2, dh2_zero, train);
  return check_launch("decoder_fwd");
}

int aae_decoder_bwd(aae_dims d, const float* dh2, const float* dec, aae_drop d1, aae_drop d2, const aae_step_state* st,
                    const float* dd1, const float* h2, float* g_d2, float* g_d1, void* stream) {
  AAE_REQUIRE(dh2 && dec && st && dd1 && h2 && g_d2 && g_d1, "null pointer");
  AAE_REQUIRE(d.B > 0 && d.H > 0 && d.C == 0 && d.D > 0 && d.H <= 2048 && d.D <= 4096, "size outside envelope");
  int ld = std::max(d.H, d.D);
  LAUNCH_R(decoder_bwd_kernel, d.B, 3 * ld, stream, d, dh2, dec, d1, d2, st, dd1, h2, g_d2, g_d1);
  return check_launch("decoder_bwd");
}

int aae_decoder_wgrad(aae_dims d, const float* inp, const float* dd1, const float* g_d2, const float* g_d1,
                      float* g_dec, aae_adam_block dec_opt, const aae_step_state* st, void* stream) {
  AAE_REQUIRE(inp && dd1 && g_d2 && g_d1 && (g_dec || dec_opt.p), "null pointer");
  AAE_REQUIRE(st || !dec_opt.p, "fused Adam needs the step state");
  const int H = d.H, D = d.D, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0; js.st = st; js.trace_id = TR_AE_WGRAD;
  OptBlock dop = opt_of(dec_opt);
  size_t off = 0;   // [Wd1 | bd1 | Wd2 | bd2]
  add_job(js, g_d1, H, inp, D, B, H, D, g_dec, dop, off); off += (size_t)H * D;
  add_job(js, g_d1, H, nullptr, 0, B, H, 1, g_dec, dop, off); off += H;
  add_job(js, g_d2, H, dd1, H, B, H, H, g_dec, dop, off); off += (size_t)H * H;
  add_job(js, g_d2, H, nullptr, 0, B, H, 1, g_dec, dop, off);
  return launch_jobs(js, as_stream(stream));
}

static int vae_dims_ok(const aae_dims& d) {
  return d.B > 0 && d.H > 0 && d.C > 0 && d.H <= 2048 && d.C + d.D <= 4096 && 2 * d.C <= 4096;
}

int aae_vae_fwd(aae_dims d, aae_bag bag, const float* h1pre, const float* cond, const float* eps, const float* enc,
                const float* dec, const aae_step_state* st, float* a1, float* mulv, float* eps_used, float* zc,
                float* h3, float* dh3_zero, double* kld_sum, void* stream) {
  AAE_REQUIRE(enc && dec && h3, "null pointer");
  AAE_REQUIRE(eps || st, "neither noise nor a step state (in-kernel Philox) given");
  AAE_REQUIRE(bag.indptr ? (bag.indices && bag.W1t && bag.v_end >= bag.v_begin) : (h1pre != nullptr),
              "neither a complete bag nor h1pre given");
  AAE_REQUIRE(d.D == 0 || cond, "condition rows missing");
  AAE_REQUIRE(vae_dims_ok(d), "size outside envelope");
  int ld = std::max(std::max(d.H, d.C + d.D), 2 * d.C);
  LAUNCH_R(vae_fwd_kernel, d.B, 2 * ld, stream, d, bag, h1pre, cond, eps, enc, dec, st, a1, mulv, eps_used, zc, h3,
           dh3_zero, kld_sum);
  return check_launch("vae_fwd");
}

int aae_vae_bwd(aae_dims d, const float* dh3, const float* enc, const float* dec, const float* eps_used, const float* a1,
                const float* mulv, const float* h3, float* g_3, float* g_ml, float* g_h1, void* stream) {
  AAE_REQUIRE(dh3 && enc && dec && eps_used && a1 && mulv && h3 && g_3 && g_ml && g_h1, "null pointer");
  AAE_REQUIRE(vae_dims_ok(d), "size outside envelope");
  int ld = std::max(std::max(d.H, d.C + d.D), 2 * d.C);
  LAUNCH_R(vae_bwd_kernel, d.B, 3 * ld, stream, d, dh3, enc, dec, eps_used, a1, mulv, h3, g_3, g_ml, g_h1);
  return check_launch("vae_bwd");
}

int aae_vae_wgrad(aae_dims d, const float* a1, const float* zc, const float* g_3, const float* g_ml, const float* g_h1,
                  float* g_enc, float* g_dec, aae_adam_block enc_opt, aae_adam_block dec_opt,
                  const aae_step_state* st, void* stream) {
  AAE_REQUIRE(a1 && zc && g_3 && g_ml && g_h1, "null pointer");
  AAE_REQUIRE((g_enc || enc_opt.p) && (g_dec || dec_opt.p), "no output");
  AAE_REQUIRE(st || (!enc_opt.p && !dec_opt.p), "fused Adam needs the step state");
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0; js.st = st; js.trace_id = TR_AE_WGRAD;
  OptBlock eo = opt_of(enc_opt), dop = opt_of(dec_opt);
  size_t off = 0;   // enc block [b1 | Wml | bml]
  add_job(js, g_h1, H, nullptr, 0, B, H, 1, g_enc, eo, off); off += H;
  add_job(js, g_ml, 2 * C, a1, H, B, 2 * C, H, g_enc, eo, off); off += (size_t)2 * C * H;
  add_job(js, g_ml, 2 * C, nullptr, 0, B, 2 * C, 1, g_enc, eo, off);
  off = 0;          // dec block [W3 | b3]
  add_job(js, g_3, H, zc, Cp, B, H, Cp, g_dec, dop, off); off += (size_t)H * Cp;
  add_job(js, g_3, H, nullptr, 0, B, H, 1, g_dec, dop, off);
  return launch_jobs(js, as_stream(stream));
}

}  // extern "C"
