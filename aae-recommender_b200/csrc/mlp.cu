// K4: the small replicated layers (encoder lin2/lin3, decoder lin1/lin2, the whole
// discriminator), forward, backward and both adversarial losses, fused into one row-local
// kernel per phase.  Rows of the batch are independent in every layer, so a CTA owns R rows and
// walks the whole chain with activations in shared memory; the (tiny, L2-resident) weights are
// streamed by every CTA.  Weight gradients (reductions over the batch) are a separate kernel.
// Reference: aaerec/aae.py:130-146 (Encoder), 165-178 (Decoder), 197-213 (Discriminator),
// 676-743 (ae_step / disc_step / gen_step), condition.py:90-99, 312-316 (concat on the code).
#include "mlp_blocks.cuh"

namespace aae {
AAE_DEFINE_TRACE_SETTER(trace_set_mlp)

// ---------------------------------------------------------------------------------------------
// ae_step forward tail.  Shared memory: [stage buffer 0 | stage buffer 1 | x (R*ld) | y (R*ld)]
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) ae_fwd_kernel(aae_dims d, aae_bag bag,
                                                             const float* __restrict__ h1pre,
                                                             const float* __restrict__ cond,
                                                             const float* __restrict__ enc,
                                                             const float* __restrict__ dec, aae_drop e1, aae_drop e2,
                                                             aae_drop d1, aae_drop d2, const aae_step_state* st,
                                                             float* a1, float* a2, float* zc, float* dd1, float* h2,
                                                             float* dh2_zero, int train) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[4];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(H, Cp);
  float* x = sm + 2 * STAGE_FLOATS;
  float* y = x + R * ld;
  float* scratch = align16f(y + R * ld);   // [MLP_THREADS/32][H]
  int row0 = blockIdx.x * R;
  if (train) trace_mark(TR_AE_FWD, 0);
  if (dh2_zero)
    for (int q = threadIdx.x; q < R * H; q += blockDim.x)
      if (row0 + q / H < B) dh2_zero[(size_t)row0 * H + q] = 0.f;
  EncBlock E(enc, H, C);
  DecBlock D(dec, H, Cp);
  aae_drop none = {nullptr, 0.f, 0};
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(E.We2, H, H);
    layers[1] = Stager::make_layer(E.We3, C, H);
    layers[2] = Stager::make_layer(D.Wd1, H, Cp);
    layers[3] = Stager::make_layer(D.Wd2, H, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 4);
  input_rows<R>(x, ld, bag, h1pre, E.b1, H, row0, B, scratch);
  drop_relu<R>(x, ld, H, row0, B, train ? e1 : none, st, a1);
  layer_fwd<R>(sg, 0, x, ld, E.be2, y, ld);
  drop_relu<R>(y, ld, H, row0, B, train ? e2 : none, st, a2);
  layer_fwd<R>(sg, 1, y, ld, E.be3, x, ld);   // z -> x[0..C)
  // concatenate the condition rows on the code (condition.py:312-316)
  for (int q = threadIdx.x; q < R * d.D; q += blockDim.x) {
    int r = q / d.D, i = q - r * d.D;
    int row = row0 + r;
    x[r * ld + C + i] = (row < B) ? cond[(size_t)row * d.D + i] : 0.f;
  }
  __syncthreads();
  if (zc) store_rows<R>(x, ld, zc, Cp, row0, B);
  layer_fwd<R>(sg, 2, x, ld, D.bd1, y, ld);
  drop_relu<R>(y, ld, H, row0, B, train ? d1 : none, st, dd1);
  layer_fwd<R>(sg, 3, y, ld, D.bd2, x, ld);
  drop_relu<R>(x, ld, H, row0, B, train ? d2 : none, st, h2);
  if (train) trace_mark(TR_AE_FWD, 1);
}

// ---------------------------------------------------------------------------------------------
// ae_step backward tail.  Shared memory: [stage 0 | stage 1 | g | t | act]
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) ae_bwd_kernel(aae_dims d, const float* __restrict__ dh2,
                                                             const float* __restrict__ enc,
                                                             const float* __restrict__ dec, aae_drop e1, aae_drop e2,
                                                             aae_drop d1, aae_drop d2, const aae_step_state* st,
                                                             const float* __restrict__ a1,
                                                             const float* __restrict__ a2,
                                                             const float* __restrict__ dd1,
                                                             const float* __restrict__ h2, float* g_d2, float* g_d1,
                                                             float* g_z, float* g_e2, float* g_h1) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[4];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(H, Cp);
  float* g = sm + 2 * STAGE_FLOATS;
  float* t = g + R * ld;
  float* act = t + R * ld;
  int row0 = blockIdx.x * R;
  trace_mark(TR_AE_BWD, 0);
  EncBlock E(enc, H, C);
  DecBlock D(dec, H, Cp);
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(D.Wd2, H, H);
    layers[1] = Stager::make_layer(D.Wd1, H, Cp);
    layers[2] = Stager::make_layer(E.We3, C, H);
    layers[3] = Stager::make_layer(E.We2, H, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 4);
  // every elementwise pass below uses the same thread -> (row, unit) mapping, so load + mask need no barrier
  load_rows<R>(g, ld, dh2, H, row0, B);
  load_rows<R>(act, ld, h2, H, row0, B);
  drop_relu_bwd<R>(g, act, ld, H, row0, B, d2, st, g_d2);
  layer_bwd<R>(sg, 0, g, ld, t, ld);
  load_rows<R>(act, ld, dd1, H, row0, B);
  drop_relu_bwd<R>(t, act, ld, H, row0, B, d1, st, g_d1);
  layer_bwd<R>(sg, 1, t, ld, g, ld);            // d(zc); only the first C entries go on
  store_rows<R>(g, ld, g_z, C, row0, B);
  layer_bwd<R>(sg, 2, g, ld, t, ld);
  load_rows<R>(act, ld, a2, H, row0, B);
  drop_relu_bwd<R>(t, act, ld, H, row0, B, e2, st, g_e2);
  layer_bwd<R>(sg, 3, t, ld, g, ld);
  load_rows<R>(act, ld, a1, H, row0, B);
  drop_relu_bwd<R>(g, act, ld, H, row0, B, e1, st, g_h1);
  trace_mark(TR_AE_BWD, 1);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// discriminator forward on one input held in zs[R][ld] (width C): layers lb (Wq1), lb+1 (Wq2), lb+2 (wq3) of
// the stager's sequence; leaves q1 in q1s, q2 in q2s, the sigmoid output per row in outs[r] (shared).
template <int R>
__device__ __forceinline__ void disc_fwd(Stager& sg, int lb, const float* zs, int ld, const DiscBlock& Q, int H,
                                         int row0, int B, const aae_drop& da, const aae_drop& db,
                                         const aae_step_state* st, float* q1s, float* q2s, float* outs) {
  layer_fwd<R>(sg, lb, zs, ld, Q.bq1, q1s, ld);
  drop_relu<R>(q1s, ld, H, row0, B, da, st, nullptr);
  layer_fwd<R>(sg, lb + 1, q1s, ld, Q.bq2, q2s, ld);
  drop_relu<R>(q2s, ld, H, row0, B, db, st, nullptr);
  layer_fwd<R>(sg, lb + 2, q2s, ld, Q.bq3, outs, 1);
  if (threadIdx.x < R) outs[threadIdx.x] = sigmoid_acc(outs[threadIdx.x]);
  __syncthreads();
}
// backward of the discriminator given g_o[r] = dL/d(lin3 pre-activation): g2 <- grad at lin2 pre-act,
// g1 <- grad at lin1 pre-act (layer lq2 of the stager's sequence is Wq2).
template <int R>
__device__ __forceinline__ void disc_bwd(Stager& sg, int lq2, const float* g_o, const DiscBlock& Q, int H, int ld,
                                         int row0, int B, const aae_drop& da, const aae_drop& db,
                                         const aae_step_state* st, const float* q1s, const float* q2s, float* g2,
                                         float* g1) {
  for (int q = threadIdx.x; q < R * H; q += blockDim.x) {
    int r = q / H, i = q - r * H;
    g2[r * ld + i] = g_o[r] * __ldg(Q.wq3 + i);
  }
  drop_relu_bwd<R>(g2, q2s, ld, H, row0, B, db, st, nullptr);   // same mapping as the loop above
  layer_bwd<R>(sg, lq2, g2, ld, g1, ld);
  drop_relu_bwd<R>(g1, q1s, ld, H, row0, B, da, st, nullptr);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// disc_step: acts row layout  [zr (C) | q1r (H) | q2r (H) | zf (C) | q1f (H) | q2f (H)]
//            grads row layout [g1r (H) | g2r (H) | gor (1) | g1f (H) | g2f (H) | gof (1)]
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) disc_phase_kernel(aae_dims d, aae_bag bag,
                                                                 const float* __restrict__ h1pre,
                                                                 const float* __restrict__ z_real, float prior_scale,
                                                                 const float* __restrict__ enc,
                                                                 const float* __restrict__ disc, aae_drop r1,
                                                                 aae_drop r2, aae_drop f1, aae_drop f2,
                                                                 const aae_step_state* st, float* acts, float* grads,
                                                                 double* loss_sum) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[10];
  const int H = d.H, C = d.C, B = d.B;
  const int ld = max(H, C);
  float* x = sm + 2 * STAGE_FLOATS;
  float* y = x + R * ld;
  float* q1 = y + R * ld;
  float* q2 = q1 + R * ld;
  float* zz = q2 + R * ld;
  float* outs = zz + R * ld;   // [R]
  float* go = outs + R;        // [R]
  float* scratch = align16f(go + R);     // [MLP_THREADS/32][H]
  int row0 = blockIdx.x * R;
  trace_mark(TR_DISC, 0);
  EncBlock E(enc, H, C);
  DiscBlock Q(disc, H, C);
  aae_drop none = {nullptr, 0.f, 0};
  const int AW = 2 * (C + 2 * H), GW = 2 * (2 * H + 1);
  // the two sides (real prior sample / encoder output) are independent until the weight gradients: they run in
  // different CTAs (blockIdx.y), which halves the dependent chain of the phase
  const int side = blockIdx.y;
  if (threadIdx.x == 0) {
    int n = 0;
    if (side) {
      layers[n++] = Stager::make_layer(E.We2, H, H);
      layers[n++] = Stager::make_layer(E.We3, C, H);
    }
    layers[n++] = Stager::make_layer(Q.Wq1, H, C);
    layers[n++] = Stager::make_layer(Q.Wq2, H, H);
    layers[n++] = Stager::make_layer(Q.wq3, 1, H);
    layers[n++] = Stager::make_layer(Q.Wq2, H, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, side ? 6 : 4);
  float lsum = 0.f;
  {
    if (side == 0) {
      // z_real ~ N(0,1) * prior_scale (aae.py:716-718)
      for (int q = threadIdx.x; q < R * C; q += blockDim.x) {
        int r = q / C, i = q - r * C;
        int row = row0 + r;
        float v = 0.f;
        if (row < B) v = z_real ? z_real[(size_t)row * C + i] : randn_elem(st, (uint32_t)(row * C + i), 77u) * prior_scale;
        zz[r * ld + i] = v;
      }
    } else {
      // z_fake = enc(batch) in eval mode (aae.py:714, 722)
      input_rows<R>(x, ld, bag, h1pre, E.b1, H, row0, B, scratch);
      drop_relu<R>(x, ld, H, row0, B, none, st, nullptr);
      layer_fwd<R>(sg, 0, x, ld, E.be2, y, ld);
      drop_relu<R>(y, ld, H, row0, B, none, st, nullptr);
      layer_fwd<R>(sg, 1, y, ld, E.be3, zz, ld);
    }
    const int lb = side ? 2 : 0;
    disc_fwd<R>(sg, lb, zz, ld, Q, H, row0, B, side ? f1 : r1, side ? f2 : r2, st, q1, q2, outs);
    if (threadIdx.x < R && row0 + threadIdx.x < B) {
      float o = outs[threadIdx.x];
      float g;
      if (side == 0) {
        float a = o + 1e-12f;                       // log(D(z_real) + TINY)
        lsum += -logf(a);
        g = (-1.0f / (float)B) / a;
      } else {
        float b = 1.0f - o + 1e-12f;                // log(1 - D(z_fake) + TINY)
        lsum += -logf(b);
        g = (1.0f / (float)B) / b;
      }
      go[threadIdx.x] = g * (1.0f - o) * o;
    }
    __syncthreads();
    // save activations, run backward, save gradients
    for (int r = 0; r < R; ++r) {
      int row = row0 + r;
      if (row >= B) break;
      float* ar = acts + (size_t)row * AW + side * (C + 2 * H);
      for (int i = threadIdx.x; i < C; i += blockDim.x) ar[i] = zz[r * ld + i];
      for (int i = threadIdx.x; i < H; i += blockDim.x) {
        ar[C + i] = q1[r * ld + i];
        ar[C + H + i] = q2[r * ld + i];
      }
    }
    disc_bwd<R>(sg, lb + 3, go, Q, H, ld, row0, B, side ? f1 : r1, side ? f2 : r2, st, q1, q2, y, x);
    for (int r = 0; r < R; ++r) {
      int row = row0 + r;
      if (row >= B) break;
      float* gr = grads + (size_t)row * GW + side * (2 * H + 1);
      for (int i = threadIdx.x; i < H; i += blockDim.x) {
        gr[i] = x[r * ld + i];
        gr[H + i] = y[r * ld + i];
      }
      if (threadIdx.x == 0) gr[2 * H] = go[r];
    }
  }
  if (threadIdx.x < R && lsum != 0.f) atomicAdd(loss_sum, (double)lsum);
  trace_mark(TR_DISC, 1);
}

// ---------------------------------------------------------------------------------------------
// gen_step
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) gen_phase_kernel(aae_dims d, aae_bag bag,
                                                                const float* __restrict__ h1pre,
                                                                const float* __restrict__ enc,
                                                                const float* __restrict__ disc, aae_drop e1,
                                                                aae_drop e2, aae_drop q1d, aae_drop q2d,
                                                                const aae_step_state* st, float* a1, float* a2,
                                                                float* g_z, float* g_e2, float* g_h1,
                                                                double* loss_sum) {
  extern __shared__ __align__(16) float sm[];
  __shared__ LayerW layers[9];
  const int H = d.H, C = d.C, B = d.B;
  const int ld = max(H, C);
  float* xa1 = sm + 2 * STAGE_FLOATS;   // a1
  float* xa2 = xa1 + R * ld;     // a2
  float* zz = xa2 + R * ld;      // z
  float* q1 = zz + R * ld;
  float* q2 = q1 + R * ld;
  float* t0 = q2 + R * ld;
  float* t1 = t0 + R * ld;
  float* outs = t1 + R * ld;
  float* go = outs + R;
  float* scratch = align16f(go + R);       // [MLP_THREADS/32][H]
  int row0 = blockIdx.x * R;
  trace_mark(TR_GEN, 0);
  EncBlock E(enc, H, C);
  DiscBlock Q(disc, H, C);
  if (threadIdx.x == 0) {
    layers[0] = Stager::make_layer(E.We2, H, H); layers[1] = Stager::make_layer(E.We3, C, H);
    layers[2] = Stager::make_layer(Q.Wq1, H, C); layers[3] = Stager::make_layer(Q.Wq2, H, H); layers[4] = Stager::make_layer(Q.wq3, 1, H);
    layers[5] = Stager::make_layer(Q.Wq2, H, H); layers[6] = Stager::make_layer(Q.Wq1, H, C); layers[7] = Stager::make_layer(E.We3, C, H); layers[8] = Stager::make_layer(E.We2, H, H);
  }
  __syncthreads();
  Stager sg;
  sg.init(sm, sm + STAGE_FLOATS, layers, 9);
  input_rows<R>(xa1, ld, bag, h1pre, E.b1, H, row0, B, scratch);
  drop_relu<R>(xa1, ld, H, row0, B, e1, st, a1);
  layer_fwd<R>(sg, 0, xa1, ld, E.be2, xa2, ld);
  drop_relu<R>(xa2, ld, H, row0, B, e2, st, a2);
  layer_fwd<R>(sg, 1, xa2, ld, E.be3, zz, ld);
  disc_fwd<R>(sg, 2, zz, ld, Q, H, row0, B, q1d, q2d, st, q1, q2, outs);
  if (threadIdx.x < R && row0 + threadIdx.x < B) {
    float o = outs[threadIdx.x];
    float a = o + 1e-12f;                           // -mean(log(D(enc(x)) + TINY)), aae.py:738
    atomicAdd(loss_sum, (double)(-logf(a)));
    go[threadIdx.x] = (-1.0f / (float)B) / a * (1.0f - o) * o;
  }
  __syncthreads();
  disc_bwd<R>(sg, 5, go, Q, H, ld, row0, B, q1d, q2d, st, q1, q2, t0, t1);   // t1 = grad at disc.lin1 pre-act
  layer_bwd<R>(sg, 6, t1, ld, t0, ld);                                      // dz
  store_rows<R>(t0, ld, g_z, C, row0, B);
  layer_bwd<R>(sg, 7, t0, ld, t1, ld);
  drop_relu_bwd<R>(t1, xa2, ld, H, row0, B, e2, st, g_e2);
  layer_bwd<R>(sg, 8, t1, ld, t0, ld);
  drop_relu_bwd<R>(t0, xa1, ld, H, row0, B, e1, st, g_h1);
  trace_mark(TR_GEN, 1);
}

}  // namespace aae

using namespace aae;

extern "C" {

static const aae_bag NO_BAG = {nullptr, nullptr, nullptr, 0, 0, 0};
static int check_bag(const aae_bag& bag, const float* h1pre) {
  if (bag.indptr) return (bag.indices && bag.W1t && bag.v_end >= bag.v_begin) ? 1 : 0;
  return h1pre ? 1 : 0;
}

int aae_ae_fwd_bag(aae_dims d, aae_bag bag, const float* h1pre, const float* cond, const float* enc, const float* dec,
                   aae_drop e1, aae_drop e2, aae_drop d1, aae_drop d2, const aae_step_state* st, float* a1, float* a2,
                   float* zc, float* dd1, float* h2, float* dh2_zero, void* stream) {
  AAE_REQUIRE(enc && dec && st && a1 && a2 && zc && dd1 && h2, "null pointer");
  AAE_REQUIRE(check_bag(bag, h1pre), "neither a complete bag nor h1pre given");
  AAE_REQUIRE(d.D == 0 || cond, "condition rows missing");
  AAE_REQUIRE(d.B > 0 && d.H > 0 && d.C > 0 && d.H <= 2048 && d.C + d.D <= 4096, "size outside envelope");
  int ld = std::max(d.H, d.C + d.D);
  LAUNCH_R(ae_fwd_kernel, d.B, 2 * ld, stream, d, bag, h1pre, cond, enc, dec, e1, e2, d1, d2, st, a1, a2, zc, dd1, h2,
           dh2_zero, 1);
  return check_launch("ae_fwd");
}

int aae_predict_tail_bag(aae_dims d, aae_bag bag, const float* h1pre, const float* cond, const float* enc,
                         const float* dec, float* h2, void* stream) {
  AAE_REQUIRE(enc && dec && h2, "null pointer");
  AAE_REQUIRE(check_bag(bag, h1pre), "neither a complete bag nor h1pre given");
  AAE_REQUIRE(d.D == 0 || cond, "condition rows missing");
  int ld = std::max(d.H, d.C + d.D);
  aae_drop none = {nullptr, 0.f, 0};
  LAUNCH_R(ae_fwd_kernel, d.B, 2 * ld, stream, d, bag, h1pre, cond, enc, dec, none, none, none, none,
           (const aae_step_state*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, h2,
           (float*)nullptr, 0);
  return check_launch("predict_tail");
}

int aae_ae_bwd(aae_dims d, const float* dh2, const float* enc, const float* dec, aae_drop e1, aae_drop e2, aae_drop d1,
               aae_drop d2, const aae_step_state* st, const float* a1, const float* a2, const float* dd1,
               const float* h2, float* g_d2, float* g_d1, float* g_z, float* g_e2, float* g_h1, void* stream) {
  AAE_REQUIRE(dh2 && enc && dec && st && a1 && a2 && dd1 && h2 && g_d2 && g_d1 && g_z && g_e2 && g_h1, "null pointer");
  int ld = std::max(d.H, d.C + d.D);
  LAUNCH_R(ae_bwd_kernel, d.B, 3 * ld, stream, d, dh2, enc, dec, e1, e2, d1, d2, st, a1, a2, dd1, h2, g_d2, g_d1, g_z,
           g_e2, g_h1);
  return check_launch("ae_bwd");
}

int aae_disc_phase_bag(aae_dims d, aae_bag bag, const float* h1pre, const float* z_real, float prior_scale,
                       const float* enc, const float* disc, aae_drop r1, aae_drop r2, aae_drop f1, aae_drop f2,
                       const aae_step_state* st, float* acts, float* grads, double* loss_sum, void* stream) {
  AAE_REQUIRE(enc && disc && st && acts && grads && loss_sum, "null pointer");
  AAE_REQUIRE(check_bag(bag, h1pre), "neither a complete bag nor h1pre given");
  int ld = std::max(d.H, d.C);
  {
    int R_ = rows_per_cta(d.B);
    size_t smem_ = sizeof(float) * ((size_t)(5 * ld + 2) * R_ + 2 * STAGE_FLOATS + SCRATCH_FLOATS(d)) + 64;
    if (R_ == 1) {
      cudaFuncSetAttribute(disc_phase_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);
      disc_phase_kernel<1><<<dim3(cdiv(d.B, 1), 2), MLP_THREADS, smem_, as_stream(stream)>>>(
          d, bag, h1pre, z_real, prior_scale, enc, disc, r1, r2, f1, f2, st, acts, grads, loss_sum);
    } else {
      cudaFuncSetAttribute(disc_phase_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);
      disc_phase_kernel<4><<<dim3(cdiv(d.B, 4), 2), MLP_THREADS, smem_, as_stream(stream)>>>(
          d, bag, h1pre, z_real, prior_scale, enc, disc, r1, r2, f1, f2, st, acts, grads, loss_sum);
    }
  }
  return check_launch("disc_phase");
}

int aae_gen_phase_bag(aae_dims d, aae_bag bag, const float* h1pre, const float* enc, const float* disc, aae_drop e1,
                      aae_drop e2, aae_drop q1, aae_drop q2, const aae_step_state* st, float* a1, float* a2, float* g_z,
                      float* g_e2, float* g_h1, double* loss_sum, void* stream) {
  AAE_REQUIRE(enc && disc && st && a1 && a2 && g_z && g_e2 && g_h1 && loss_sum, "null pointer");
  AAE_REQUIRE(check_bag(bag, h1pre), "neither a complete bag nor h1pre given");
  int ld = std::max(d.H, d.C);
  LAUNCH_R(gen_phase_kernel, d.B, 7 * ld + 2, stream, d, bag, h1pre, enc, disc, e1, e2, q1, q2, st, a1, a2, g_z, g_e2,
           g_h1, loss_sum);
  return check_launch("gen_phase");
}

int aae_ae_wgrad(aae_dims d, const float* a1, const float* a2, const float* zc, const float* dd1, const float* g_d2,
                 const float* g_d1, const float* g_z, const float* g_e2, const float* g_h1, float* g_enc, float* g_dec,
                 aae_adam_block enc_opt, aae_adam_block dec_opt, const aae_step_state* st, void* stream) {
  AAE_REQUIRE(a1 && a2 && zc && dd1 && g_d2 && g_d1 && g_z && g_e2 && g_h1, "null pointer");
  AAE_REQUIRE((g_enc || enc_opt.p) && (g_dec || dec_opt.p), "no output");
  AAE_REQUIRE(st || (!enc_opt.p && !dec_opt.p), "fused Adam needs the step state");
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0; js.st = st; js.trace_id = TR_AE_WGRAD;
  // enc block [b1 | We2 | be2 | We3 | be3]
  OptBlock eo = opt_of(enc_opt), dop = opt_of(dec_opt);
  size_t off = 0;
  add_job(js, g_h1, H, nullptr, 0, B, H, 1, g_enc, eo, off); off += H;
  add_job(js, g_e2, H, a1, H, B, H, H, g_enc, eo, off); off += (size_t)H * H;
  add_job(js, g_e2, H, nullptr, 0, B, H, 1, g_enc, eo, off); off += H;
  add_job(js, g_z, C, a2, H, B, C, H, g_enc, eo, off); off += (size_t)C * H;
  add_job(js, g_z, C, nullptr, 0, B, C, 1, g_enc, eo, off);
  // dec block [Wd1 | bd1 | Wd2 | bd2]
  off = 0;
  add_job(js, g_d1, H, zc, Cp, B, H, Cp, g_dec, dop, off); off += (size_t)H * Cp;
  add_job(js, g_d1, H, nullptr, 0, B, H, 1, g_dec, dop, off); off += H;
  add_job(js, g_d2, H, dd1, H, B, H, H, g_dec, dop, off); off += (size_t)H * H;
  add_job(js, g_d2, H, nullptr, 0, B, H, 1, g_dec, dop, off);
  return launch_jobs(js, as_stream(stream));
}

int aae_disc_wgrad(aae_dims d, const float* acts, const float* grads, float* g_disc, aae_adam_block disc_opt,
                   const aae_step_state* st, void* stream) {
  AAE_REQUIRE(acts && grads && (g_disc || disc_opt.p), "null pointer");
  AAE_REQUIRE(st || !disc_opt.p, "fused Adam needs the step state");
  const int H = d.H, C = d.C, B = d.B;
  // two virtual rows (real, fake) per batch row
  const int AW = C + 2 * H, GW = 2 * H + 1;
  WJobs js;
  js.n = 0; js.total = 0; js.st = st; js.trace_id = TR_DISC_WGRAD;
  OptBlock qo = opt_of(disc_opt);
  size_t off = 0;  // [Wq1 | bq1 | Wq2 | bq2 | wq3 | bq3]
  add_job(js, grads, GW, acts, AW, 2 * B, H, C, g_disc, qo, off); off += (size_t)H * C;
  add_job(js, grads, GW, nullptr, 0, 2 * B, H, 1, g_disc, qo, off); off += H;
  add_job(js, grads + H, GW, acts + C, AW, 2 * B, H, H, g_disc, qo, off); off += (size_t)H * H;
  add_job(js, grads + H, GW, nullptr, 0, 2 * B, H, 1, g_disc, qo, off); off += H;
  add_job(js, grads + 2 * H, GW, acts + C + H, AW, 2 * B, 1, H, g_disc, qo, off); off += H;
  add_job(js, grads + 2 * H, GW, nullptr, 0, 2 * B, 1, 1, g_disc, qo, off);
  return launch_jobs(js, as_stream(stream));
}

int aae_gen_wgrad(aae_dims d, const float* a1, const float* a2, const float* g_z, const float* g_e2, const float* g_h1,
                  float* g_enc, aae_adam_block enc_opt, const aae_step_state* st, void* stream) {
  AAE_REQUIRE(a1 && a2 && g_z && g_e2 && g_h1 && (g_enc || enc_opt.p), "null pointer");
  AAE_REQUIRE(st || !enc_opt.p, "fused Adam needs the step state");
  const int H = d.H, C = d.C, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0; js.st = st; js.trace_id = TR_GEN_WGRAD;
  OptBlock eo = opt_of(enc_opt);
  size_t off = 0;
  add_job(js, g_h1, H, nullptr, 0, B, H, 1, g_enc, eo, off); off += H;
  add_job(js, g_e2, H, a1, H, B, H, H, g_enc, eo, off); off += (size_t)H * H;
  add_job(js, g_e2, H, nullptr, 0, B, H, 1, g_enc, eo, off); off += H;
  add_job(js, g_z, C, a2, H, B, C, H, g_enc, eo, off); off += (size_t)C * H;
  add_job(js, g_z, C, nullptr, 0, B, C, 1, g_enc, eo, off);
  return launch_jobs(js, as_stream(stream));
}

}  // extern "C"
