// K4: the small replicated layers (encoder lin2/lin3, decoder lin1/lin2, the whole
// discriminator), forward, backward and both adversarial losses, fused into one row-local
// kernel per phase.  Rows of the batch are independent in every layer, so a CTA owns R rows and
// walks the whole chain with activations in shared memory; the (tiny, L2-resident) weights are
// streamed by every CTA.  Weight gradients (reductions over the batch) are a separate kernel.
// Reference: aaerec/aae.py:130-146 (Encoder), 165-178 (Decoder), 197-213 (Discriminator),
// 676-743 (ae_step / disc_step / gen_step), condition.py:90-99, 312-316 (concat on the code).
#include "common.cuh"

namespace aae {

constexpr int MLP_THREADS = 128;

// y[r][o] = b[o] + sum_i x[r][i] * W[o*I + i]; one warp per output, lanes over i (coalesced rows).
template <int R>
__device__ __forceinline__ void row_linear(const float* xs, int ldx, int I, const float* __restrict__ W,
                                           const float* __restrict__ b, int O, float* ys, int ldy) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o0 = warp * 2; o0 < O; o0 += nw * 2) {
    int o1 = o0 + 1;
    bool has1 = o1 < O;
    float acc0[R], acc1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc0[r] = acc1[r] = 0.f;
    const float* w0 = W + (size_t)o0 * I;
    const float* w1 = W + (size_t)(has1 ? o1 : o0) * I;
    for (int i = lane; i < I; i += 32) {
      float a = __ldg(w0 + i), c = __ldg(w1 + i);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float x = xs[r * ldx + i];
        acc0[r] = fmaf(a, x, acc0[r]);
        acc1[r] = fmaf(c, x, acc1[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s0 = warp_sum(acc0[r]), s1 = warp_sum(acc1[r]);
      if (lane == 0) {
        ys[r * ldy + o0] = s0 + b[o0];
        if (has1) ys[r * ldy + o1] = s1 + b[o1];
      }
    }
  }
}

// dx[r][i] = sum_o dy[r][o] * W[o*I + i]; one thread per i (coalesced over i).
template <int R>
__device__ __forceinline__ void row_linear_bwd(const float* dys, int ldy, int O, const float* __restrict__ W, int I,
                                               float* dxs, int ldx) {
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    int o = 0;
    for (; o + 4 <= O; o += 4) {
      float w0 = __ldg(W + (size_t)o * I + i), w1 = __ldg(W + (size_t)(o + 1) * I + i);
      float w2 = __ldg(W + (size_t)(o + 2) * I + i), w3 = __ldg(W + (size_t)(o + 3) * I + i);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* d = dys + r * ldy + o;
        acc[r] = fmaf(w0, d[0], acc[r]);
        acc[r] = fmaf(w1, d[1], acc[r]);
        acc[r] = fmaf(w2, d[2], acc[r]);
        acc[r] = fmaf(w3, d[3], acc[r]);
      }
    }
    for (; o < O; ++o) {
      float w0 = __ldg(W + (size_t)o * I + i);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(w0, dys[r * ldy + o], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) dxs[r * ldx + i] = acc[r];
  }
}

// in place: x <- relu(x * dropfactor); optionally mirrored to global
template <int R>
__device__ __forceinline__ void drop_relu(float* xs, int ld, int n, int row0, int B, const aae_drop& d,
                                          const aae_step_state* st, float* gout) {
  for (int q = threadIdx.x; q < R * n; q += blockDim.x) {
    int r = q / n, i = q - r * n;
    int row = row0 + r;
    if (row >= B) continue;
    float f = drop_factor(d, st, (uint32_t)(row * n + i));
    float v = fmaxf(xs[r * ld + i] * f, 0.f);
    xs[r * ld + i] = v;
    if (gout) gout[(size_t)row * n + i] = v;
  }
}
// in place: g <- g * 1[act > 0] * dropfactor; mirrored to global
template <int R>
__device__ __forceinline__ void drop_relu_bwd(float* gs, const float* acts, int ld, int n, int row0, int B,
                                              const aae_drop& d, const aae_step_state* st, float* gout) {
  for (int q = threadIdx.x; q < R * n; q += blockDim.x) {
    int r = q / n, i = q - r * n;
    int row = row0 + r;
    if (row >= B) continue;
    float f = drop_factor(d, st, (uint32_t)(row * n + i));
    float v = (acts[r * ld + i] > 0.f) ? gs[r * ld + i] * f : 0.f;
    gs[r * ld + i] = v;
    if (gout) gout[(size_t)row * n + i] = v;
  }
}
template <int R>
__device__ __forceinline__ void load_rows(float* xs, int ld, const float* g, int n, int row0, int B) {
  for (int q = threadIdx.x; q < R * n; q += blockDim.x) {
    int r = q / n, i = q - r * n;
    int row = row0 + r;
    xs[r * ld + i] = (row < B) ? g[(size_t)row * n + i] : 0.f;
  }
}
template <int R>
__device__ __forceinline__ void store_rows(const float* xs, int ld, float* g, int n, int row0, int B) {
  for (int q = threadIdx.x; q < R * n; q += blockDim.x) {
    int r = q / n, i = q - r * n;
    int row = row0 + r;
    if (row < B) g[(size_t)row * n + i] = xs[r * ld + i];
  }
}

struct EncBlock {
  const float *b1, *We2, *be2, *We3, *be3;
  __device__ EncBlock(const float* p, int H, int C) {
    b1 = p; We2 = b1 + H; be2 = We2 + (size_t)H * H; We3 = be2 + H; be3 = We3 + (size_t)C * H;
  }
};
struct DecBlock {
  const float *Wd1, *bd1, *Wd2, *bd2;
  __device__ DecBlock(const float* p, int H, int Cp) {
    Wd1 = p; bd1 = Wd1 + (size_t)H * Cp; Wd2 = bd1 + H; bd2 = Wd2 + (size_t)H * H;
  }
};
struct DiscBlock {
  const float *Wq1, *bq1, *Wq2, *bq2, *wq3, *bq3;
  __device__ DiscBlock(const float* p, int H, int C) {
    Wq1 = p; bq1 = Wq1 + (size_t)H * C; Wq2 = bq1 + H; bq2 = Wq2 + (size_t)H * H; wq3 = bq2 + H; bq3 = wq3 + H;
  }
};

// ---------------------------------------------------------------------------------------------
// ae_step forward tail
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) ae_fwd_kernel(aae_dims d, const float* __restrict__ h1pre,
                                                             const float* __restrict__ cond,
                                                             const float* __restrict__ enc,
                                                             const float* __restrict__ dec, aae_drop e1, aae_drop e2,
                                                             aae_drop d1, aae_drop d2, const aae_step_state* st,
                                                             float* a1, float* a2, float* zc, float* dd1, float* h2,
                                                             int train) {
  extern __shared__ float sm[];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(H, Cp);
  float* x = sm;
  float* y = sm + R * ld;
  int row0 = blockIdx.x * R;
  EncBlock E(enc, H, C);
  DecBlock D(dec, H, Cp);
  aae_drop none = {nullptr, 0.f, 0};
  load_rows<R>(x, ld, h1pre, H, row0, B);
  __syncthreads();
  drop_relu<R>(x, ld, H, row0, B, train ? e1 : none, st, a1);
  __syncthreads();
  row_linear<R>(x, ld, H, E.We2, E.be2, H, y, ld);
  __syncthreads();
  drop_relu<R>(y, ld, H, row0, B, train ? e2 : none, st, a2);
  __syncthreads();
  row_linear<R>(y, ld, H, E.We3, E.be3, C, x, ld);   // z -> x[0..C)
  // concatenate the condition rows on the code (condition.py:312-316)
  for (int q = threadIdx.x; q < R * d.D; q += blockDim.x) {
    int r = q / d.D, i = q - r * d.D;
    int row = row0 + r;
    x[r * ld + C + i] = (row < B) ? cond[(size_t)row * d.D + i] : 0.f;
  }
  __syncthreads();
  if (zc) store_rows<R>(x, ld, zc, Cp, row0, B);
  row_linear<R>(x, ld, Cp, D.Wd1, D.bd1, H, y, ld);
  __syncthreads();
  drop_relu<R>(y, ld, H, row0, B, train ? d1 : none, st, dd1);
  __syncthreads();
  row_linear<R>(y, ld, H, D.Wd2, D.bd2, H, x, ld);
  __syncthreads();
  drop_relu<R>(x, ld, H, row0, B, train ? d2 : none, st, h2);
}

// ---------------------------------------------------------------------------------------------
// ae_step backward tail
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
  extern __shared__ float sm[];
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  const int ld = max(H, Cp);
  float* g = sm;
  float* t = sm + R * ld;
  float* act = sm + 2 * R * ld;
  int row0 = blockIdx.x * R;
  EncBlock E(enc, H, C);
  DecBlock D(dec, H, Cp);
  load_rows<R>(g, ld, dh2, H, row0, B);
  load_rows<R>(act, ld, h2, H, row0, B);
  __syncthreads();
  drop_relu_bwd<R>(g, act, ld, H, row0, B, d2, st, g_d2);
  __syncthreads();
  row_linear_bwd<R>(g, ld, H, D.Wd2, H, t, ld);
  load_rows<R>(act, ld, dd1, H, row0, B);
  __syncthreads();
  drop_relu_bwd<R>(t, act, ld, H, row0, B, d1, st, g_d1);
  __syncthreads();
  row_linear_bwd<R>(t, ld, H, D.Wd1, Cp, g, ld);    // d(zc); only the first C entries go on
  __syncthreads();
  store_rows<R>(g, ld, g_z, C, row0, B);
  row_linear_bwd<R>(g, ld, C, E.We3, H, t, ld);
  load_rows<R>(act, ld, a2, H, row0, B);
  __syncthreads();
  drop_relu_bwd<R>(t, act, ld, H, row0, B, e2, st, g_e2);
  __syncthreads();
  row_linear_bwd<R>(t, ld, H, E.We2, H, g, ld);
  load_rows<R>(act, ld, a1, H, row0, B);
  __syncthreads();
  drop_relu_bwd<R>(g, act, ld, H, row0, B, e1, st, g_h1);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// discriminator forward on one input held in zs[R][ld] (width C); leaves q1 in q1s, q2 in q2s,
// returns sigmoid output per row in outs[r] (shared).
template <int R>
__device__ __forceinline__ void disc_fwd(const float* zs, int ld, const DiscBlock& Q, int H, int C, int row0, int B,
                                         const aae_drop& da, const aae_drop& db, const aae_step_state* st,
                                         float* q1s, float* q2s, float* outs) {
  row_linear<R>(zs, ld, C, Q.Wq1, Q.bq1, H, q1s, ld);
  __syncthreads();
  drop_relu<R>(q1s, ld, H, row0, B, da, st, nullptr);
  __syncthreads();
  row_linear<R>(q1s, ld, H, Q.Wq2, Q.bq2, H, q2s, ld);
  __syncthreads();
  drop_relu<R>(q2s, ld, H, row0, B, db, st, nullptr);
  __syncthreads();
  row_linear<R>(q2s, ld, H, Q.wq3, Q.bq3, 1, outs, 1);
  __syncthreads();
  if (threadIdx.x < R) outs[threadIdx.x] = sigmoid_acc(outs[threadIdx.x]);
  __syncthreads();
}
// backward of the discriminator given g_o[r] = dL/d(lin3 pre-activation): g2 <- grad at lin2 pre-act,
// g1 <- grad at lin1 pre-act.
template <int R>
__device__ __forceinline__ void disc_bwd(const float* g_o, const DiscBlock& Q, int H, int ld, int row0, int B,
                                         const aae_drop& da, const aae_drop& db, const aae_step_state* st,
                                         const float* q1s, const float* q2s, float* g2, float* g1) {
  for (int q = threadIdx.x; q < R * H; q += blockDim.x) {
    int r = q / H, i = q - r * H;
    g2[r * ld + i] = g_o[r] * __ldg(Q.wq3 + i);
  }
  __syncthreads();
  drop_relu_bwd<R>(g2, q2s, ld, H, row0, B, db, st, nullptr);
  __syncthreads();
  row_linear_bwd<R>(g2, ld, H, Q.Wq2, H, g1, ld);
  __syncthreads();
  drop_relu_bwd<R>(g1, q1s, ld, H, row0, B, da, st, nullptr);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// disc_step: acts row layout  [zr (C) | q1r (H) | q2r (H) | zf (C) | q1f (H) | q2f (H)]
//            grads row layout [g1r (H) | g2r (H) | gor (1) | g1f (H) | g2f (H) | gof (1)]
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) disc_phase_kernel(aae_dims d, const float* __restrict__ h1pre,
                                                                 const float* __restrict__ z_real, float prior_scale,
                                                                 const float* __restrict__ enc,
                                                                 const float* __restrict__ disc, aae_drop r1,
                                                                 aae_drop r2, aae_drop f1, aae_drop f2,
                                                                 const aae_step_state* st, float* acts, float* grads,
                                                                 double* loss_sum) {
  extern __shared__ float sm[];
  const int H = d.H, C = d.C, B = d.B;
  const int ld = max(H, C);
  float* x = sm;
  float* y = x + R * ld;
  float* q1 = y + R * ld;
  float* q2 = q1 + R * ld;
  float* zz = q2 + R * ld;
  float* outs = zz + R * ld;   // [R]
  float* go = outs + R;        // [R]
  int row0 = blockIdx.x * R;
  EncBlock E(enc, H, C);
  DiscBlock Q(disc, H, C);
  aae_drop none = {nullptr, 0.f, 0};
  const int AW = 2 * (C + 2 * H), GW = 2 * (2 * H + 1);
  float lsum = 0.f;
  for (int side = 0; side < 2; ++side) {
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
      load_rows<R>(x, ld, h1pre, H, row0, B);
      __syncthreads();
      drop_relu<R>(x, ld, H, row0, B, none, st, nullptr);
      __syncthreads();
      row_linear<R>(x, ld, H, E.We2, E.be2, H, y, ld);
      __syncthreads();
      drop_relu<R>(y, ld, H, row0, B, none, st, nullptr);
      __syncthreads();
      row_linear<R>(y, ld, H, E.We3, E.be3, C, zz, ld);
    }
    __syncthreads();
    disc_fwd<R>(zz, ld, Q, H, C, row0, B, side ? f1 : r1, side ? f2 : r2, st, q1, q2, outs);
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
    disc_bwd<R>(go, Q, H, ld, row0, B, side ? f1 : r1, side ? f2 : r2, st, q1, q2, y, x);
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
    __syncthreads();
  }
  if (threadIdx.x < R && lsum != 0.f) atomicAdd(loss_sum, (double)lsum);
}

// ---------------------------------------------------------------------------------------------
// gen_step
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(MLP_THREADS) gen_phase_kernel(aae_dims d, const float* __restrict__ h1pre,
                                                                const float* __restrict__ enc,
                                                                const float* __restrict__ disc, aae_drop e1,
                                                                aae_drop e2, aae_drop q1d, aae_drop q2d,
                                                                const aae_step_state* st, float* a1, float* a2,
                                                                float* g_z, float* g_e2, float* g_h1,
                                                                double* loss_sum) {
  extern __shared__ float sm[];
  const int H = d.H, C = d.C, B = d.B;
  const int ld = max(H, C);
  float* xa1 = sm;               // a1
  float* xa2 = xa1 + R * ld;     // a2
  float* zz = xa2 + R * ld;      // z
  float* q1 = zz + R * ld;
  float* q2 = q1 + R * ld;
  float* t0 = q2 + R * ld;
  float* t1 = t0 + R * ld;
  float* outs = t1 + R * ld;
  float* go = outs + R;
  int row0 = blockIdx.x * R;
  EncBlock E(enc, H, C);
  DiscBlock Q(disc, H, C);
  load_rows<R>(xa1, ld, h1pre, H, row0, B);
  __syncthreads();
  drop_relu<R>(xa1, ld, H, row0, B, e1, st, a1);
  __syncthreads();
  row_linear<R>(xa1, ld, H, E.We2, E.be2, H, xa2, ld);
  __syncthreads();
  drop_relu<R>(xa2, ld, H, row0, B, e2, st, a2);
  __syncthreads();
  row_linear<R>(xa2, ld, H, E.We3, E.be3, C, zz, ld);
  __syncthreads();
  disc_fwd<R>(zz, ld, Q, H, C, row0, B, q1d, q2d, st, q1, q2, outs);
  if (threadIdx.x < R && row0 + threadIdx.x < B) {
    float o = outs[threadIdx.x];
    float a = o + 1e-12f;                           // -mean(log(D(enc(x)) + TINY)), aae.py:738
    atomicAdd(loss_sum, (double)(-logf(a)));
    go[threadIdx.x] = (-1.0f / (float)B) / a * (1.0f - o) * o;
  }
  __syncthreads();
  disc_bwd<R>(go, Q, H, ld, row0, B, q1d, q2d, st, q1, q2, t0, t1);   // t1 = grad at disc.lin1 pre-act
  row_linear_bwd<R>(t1, ld, H, Q.Wq1, C, t0, ld);                    // dz
  __syncthreads();
  store_rows<R>(t0, ld, g_z, C, row0, B);
  row_linear_bwd<R>(t0, ld, C, E.We3, H, t1, ld);
  __syncthreads();
  drop_relu_bwd<R>(t1, xa2, ld, H, row0, B, e2, st, g_e2);
  __syncthreads();
  row_linear_bwd<R>(t1, ld, H, E.We2, H, t0, ld);
  __syncthreads();
  drop_relu_bwd<R>(t0, xa1, ld, H, row0, B, e1, st, g_h1);
}

// ---------------------------------------------------------------------------------------------
// weight gradients of the small layers: dW[o,i] = sum_r dY[r,o] * X[r,i], db[o] = sum_r dY[r,o].
// Up to 8 jobs per launch; one thread per output element (i fastest -> coalesced X reads).
// ---------------------------------------------------------------------------------------------
struct WJob {
  const float* dY; int ldy;   // row pitch
  const float* X;  int ldx;   // X == nullptr -> bias job (X == 1)
  int rows, O, I;
  float* out;                 // [O, I]
  int begin;                  // first linear output index of this job
};
struct WJobs {
  WJob j[10];
  int n, total;
};
__global__ void __launch_bounds__(256) small_wgrad_kernel(WJobs jobs) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= jobs.total) return;
  int k = 0;
#pragma unroll
  for (int q = 1; q < 10; ++q)
    if (q < jobs.n && idx >= jobs.j[q].begin) k = q;
  const WJob& J = jobs.j[k];
  int e = idx - J.begin;
  int o = e / J.I, i = e - o * J.I;
  float acc0 = 0.f, acc1 = 0.f;
  int r = 0;
  if (J.X) {
    for (; r + 2 <= J.rows; r += 2) {
      acc0 = fmaf(J.dY[(size_t)r * J.ldy + o], J.X[(size_t)r * J.ldx + i], acc0);
      acc1 = fmaf(J.dY[(size_t)(r + 1) * J.ldy + o], J.X[(size_t)(r + 1) * J.ldx + i], acc1);
    }
    for (; r < J.rows; ++r) acc0 = fmaf(J.dY[(size_t)r * J.ldy + o], J.X[(size_t)r * J.ldx + i], acc0);
  } else {
    for (; r < J.rows; ++r) acc0 += J.dY[(size_t)r * J.ldy + o];
  }
  J.out[e] = acc0 + acc1;
}

static void add_job(WJobs& js, const float* dY, int ldy, const float* X, int ldx, int rows, int O, int I, float* out) {
  WJob& j = js.j[js.n++];
  j.dY = dY; j.ldy = ldy; j.X = X; j.ldx = ldx; j.rows = rows; j.O = O; j.I = I; j.out = out;
  j.begin = js.total;
  js.total += O * I;
}
static int launch_jobs(const WJobs& js, cudaStream_t s) {
  small_wgrad_kernel<<<cdiv(js.total, 256), 256, 0, s>>>(js);
  return check_launch("small_wgrad");
}

static inline int rows_per_cta(int B) { return B > 2048 ? 4 : 1; }

}  // namespace aae

using namespace aae;

#define LAUNCH_R(kernel, B, smem_floats_per_row, stream, ...)                                          \
  do {                                                                                                 \
    int R_ = rows_per_cta(B);                                                                          \
    size_t smem_ = sizeof(float) * (size_t)(smem_floats_per_row) * R_ + 64;                            \
    if (R_ == 1)                                                                                       \
      kernel<1><<<cdiv(B, 1), MLP_THREADS, smem_, as_stream(stream)>>>(__VA_ARGS__);                   \
    else                                                                                               \
      kernel<4><<<cdiv(B, 4), MLP_THREADS, smem_, as_stream(stream)>>>(__VA_ARGS__);                   \
  } while (0)

extern "C" {

int aae_ae_fwd(aae_dims d, const float* h1pre, const float* cond, const float* enc, const float* dec, aae_drop e1,
               aae_drop e2, aae_drop d1, aae_drop d2, const aae_step_state* st, float* a1, float* a2, float* zc,
               float* dd1, float* h2, void* stream) {
  AAE_REQUIRE(h1pre && enc && dec && st && a1 && a2 && zc && dd1 && h2, "null pointer");
  AAE_REQUIRE(d.D == 0 || cond, "condition rows missing");
  AAE_REQUIRE(d.B > 0 && d.H > 0 && d.C > 0 && d.H <= 2048 && d.C + d.D <= 4096, "size outside envelope");
  int ld = std::max(d.H, d.C + d.D);
  LAUNCH_R(ae_fwd_kernel, d.B, 2 * ld, stream, d, h1pre, cond, enc, dec, e1, e2, d1, d2, st, a1, a2, zc, dd1, h2, 1);
  return check_launch("ae_fwd");
}

int aae_predict_tail(aae_dims d, const float* h1pre, const float* cond, const float* enc, const float* dec, float* h2,
                     void* stream) {
  AAE_REQUIRE(h1pre && enc && dec && h2, "null pointer");
  AAE_REQUIRE(d.D == 0 || cond, "condition rows missing");
  int ld = std::max(d.H, d.C + d.D);
  aae_drop none = {nullptr, 0.f, 0};
  LAUNCH_R(ae_fwd_kernel, d.B, 2 * ld, stream, d, h1pre, cond, enc, dec, none, none, none, none,
           (const aae_step_state*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, h2, 0);
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

int aae_disc_phase(aae_dims d, const float* h1pre, const float* z_real, float prior_scale, const float* enc,
                   const float* disc, aae_drop r1, aae_drop r2, aae_drop f1, aae_drop f2, const aae_step_state* st,
                   float* acts, float* grads, double* loss_sum, void* stream) {
  AAE_REQUIRE(h1pre && enc && disc && st && acts && grads && loss_sum, "null pointer");
  int ld = std::max(d.H, d.C);
  LAUNCH_R(disc_phase_kernel, d.B, 5 * ld + 2, stream, d, h1pre, z_real, prior_scale, enc, disc, r1, r2, f1, f2, st,
           acts, grads, loss_sum);
  return check_launch("disc_phase");
}

int aae_gen_phase(aae_dims d, const float* h1pre, const float* enc, const float* disc, aae_drop e1, aae_drop e2,
                  aae_drop q1, aae_drop q2, const aae_step_state* st, float* a1, float* a2, float* g_z, float* g_e2,
                  float* g_h1, double* loss_sum, void* stream) {
  AAE_REQUIRE(h1pre && enc && disc && st && a1 && a2 && g_z && g_e2 && g_h1 && loss_sum, "null pointer");
  int ld = std::max(d.H, d.C);
  LAUNCH_R(gen_phase_kernel, d.B, 7 * ld + 2, stream, d, h1pre, enc, disc, e1, e2, q1, q2, st, a1, a2, g_z, g_e2, g_h1,
           loss_sum);
  return check_launch("gen_phase");
}

int aae_ae_wgrad(aae_dims d, const float* a1, const float* a2, const float* zc, const float* dd1, const float* g_d2,
                 const float* g_d1, const float* g_z, const float* g_e2, const float* g_h1, float* g_enc, float* g_dec,
                 void* stream) {
  AAE_REQUIRE(a1 && a2 && zc && dd1 && g_d2 && g_d1 && g_z && g_e2 && g_h1 && g_enc && g_dec, "null pointer");
  const int H = d.H, C = d.C, Cp = d.C + d.D, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0;
  // enc block [b1 | We2 | be2 | We3 | be3]
  float* p = g_enc;
  add_job(js, g_h1, H, nullptr, 0, B, H, 1, p); p += H;
  add_job(js, g_e2, H, a1, H, B, H, H, p); p += (size_t)H * H;
  add_job(js, g_e2, H, nullptr, 0, B, H, 1, p); p += H;
  add_job(js, g_z, C, a2, H, B, C, H, p); p += (size_t)C * H;
  add_job(js, g_z, C, nullptr, 0, B, C, 1, p);
  // dec block [Wd1 | bd1 | Wd2 | bd2]
  p = g_dec;
  add_job(js, g_d1, H, zc, Cp, B, H, Cp, p); p += (size_t)H * Cp;
  add_job(js, g_d1, H, nullptr, 0, B, H, 1, p); p += H;
  add_job(js, g_d2, H, dd1, H, B, H, H, p); p += (size_t)H * H;
  add_job(js, g_d2, H, nullptr, 0, B, H, 1, p);
  return launch_jobs(js, as_stream(stream));
}

int aae_disc_wgrad(aae_dims d, const float* acts, const float* grads, float* g_disc, void* stream) {
  AAE_REQUIRE(acts && grads && g_disc, "null pointer");
  const int H = d.H, C = d.C, B = d.B;
  // two virtual rows (real, fake) per batch row
  const int AW = C + 2 * H, GW = 2 * H + 1;
  WJobs js;
  js.n = 0; js.total = 0;
  float* p = g_disc;  // [Wq1 | bq1 | Wq2 | bq2 | wq3 | bq3]
  add_job(js, grads, GW, acts, AW, 2 * B, H, C, p); p += (size_t)H * C;
  add_job(js, grads, GW, nullptr, 0, 2 * B, H, 1, p); p += H;
  add_job(js, grads + H, GW, acts + C, AW, 2 * B, H, H, p); p += (size_t)H * H;
  add_job(js, grads + H, GW, nullptr, 0, 2 * B, H, 1, p); p += H;
  add_job(js, grads + 2 * H, GW, acts + C + H, AW, 2 * B, 1, H, p); p += H;
  add_job(js, grads + 2 * H, GW, nullptr, 0, 2 * B, 1, 1, p);
  return launch_jobs(js, as_stream(stream));
}

int aae_gen_wgrad(aae_dims d, const float* a1, const float* a2, const float* g_z, const float* g_e2, const float* g_h1,
                  float* g_enc, void* stream) {
  AAE_REQUIRE(a1 && a2 && g_z && g_e2 && g_h1 && g_enc, "null pointer");
  const int H = d.H, C = d.C, B = d.B;
  WJobs js;
  js.n = 0; js.total = 0;
  float* p = g_enc;
  add_job(js, g_h1, H, nullptr, 0, B, H, 1, p); p += H;
  add_job(js, g_e2, H, a1, H, B, H, H, p); p += (size_t)H * H;
  add_job(js, g_e2, H, nullptr, 0, B, H, 1, p); p += H;
  add_job(js, g_z, C, a2, H, B, C, H, p); p += (size_t)C * H;
  add_job(js, g_z, C, nullptr, 0, B, C, 1, p);
  return launch_jobs(js, as_stream(stream));
}

}  // extern "C"
