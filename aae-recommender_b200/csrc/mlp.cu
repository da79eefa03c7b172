// K4: the small replicated layers (encoder lin2/lin3, decoder lin1/lin2, the whole
// discriminator), forward, backward and both adversarial losses, fused into one row-local
// kernel per phase.  Rows of the batch are independent in every layer, so a CTA owns R rows and
// walks the whole chain with activations in shared memory; the (tiny, L2-resident) weights are
// streamed by every CTA.  Weight gradients (reductions over the batch) are a separate kernel.
// Reference: aaerec/aae.py:130-146 (Encoder), 165-178 (Decoder), 197-213 (Discriminator),
// 676-743 (ae_step / disc_step / gen_step), condition.py:90-99, 312-316 (concat on the code).
#include "common.cuh"

namespace aae {
AAE_DEFINE_TRACE_SETTER(trace_set_mlp)

constexpr int MLP_THREADS = 256;
constexpr int STAGE_FLOATS = 10240;   // one staging buffer (40 KB); two of them per CTA

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One weight matrix of a kernel's layer sequence, torch layout [O, I] row-major.
struct LayerW {
  const float* W;
  int O, I;
  int rpc;   // rows per staged chunk (filled by make_layer)
};
// Rows [r0, r1) of layer `layer`, resident in shared memory at w (row pitch I).
struct Chunk {
  const float* w;
  int layer, r0, r1;
};

// Streams the weight matrices of the kernel's whole layer sequence through two shared-memory buffers with
// cp.async, one chunk of whole rows at a time, always one chunk ahead of the consumer (across layer
// boundaries too: the weights do not depend on the activations).  All threads call every method.
struct Stager {
  float *buf0, *buf1;
  const LayerW* L;
  int n;
  int pl, pr, pbuf;   // next chunk to prefetch
  int cl, cr, cbuf;   // next chunk to consume
  int inflight;

  __device__ static int rows_per_chunk(int I) {
    int a = (I & 3) == 0 ? 1 : ((I & 1) == 0 ? 2 : 4);   // chunk starts stay 16-byte aligned
    int r = STAGE_FLOATS / I;
    if (r >= a) r -= r % a;
    return max(r, 1);
  }
  __device__ static LayerW make_layer(const float* W, int O, int I) {
    LayerW l;
    l.W = W; l.O = O; l.I = I; l.rpc = rows_per_chunk(I);
    return l;
  }
  __device__ void init(float* b0, float* b1, const LayerW* layers, int nlayers) {
    buf0 = b0; buf1 = b1; L = layers; n = nlayers;
    pl = pr = pbuf = cl = cr = cbuf = inflight = 0;
    issue();
  }
  __device__ void issue() {
    if (pl >= n) return;
    const LayerW l = L[pl];
    const int rc = min(l.rpc, l.O - pr);
    const float* src = l.W + (size_t)pr * l.I;
    float* dst = pbuf ? buf1 : buf0;
    const int nf = rc * l.I;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int n4 = nf >> 2;
      for (int q = threadIdx.x; q < n4; q += blockDim.x) cp_async16(dst + 4 * q, src + 4 * q);
      for (int q = (n4 << 2) + threadIdx.x; q < nf; q += blockDim.x) cp_async4(dst + q, src + q);
    } else {
      for (int q = threadIdx.x; q < nf; q += blockDim.x) cp_async4(dst + q, src + q);
    }
    cp_async_commit();
    pbuf ^= 1;
    ++inflight;
    pr += rc;
    if (pr >= l.O) { ++pl; pr = 0; }
  }
  // Next chunk, ready in shared memory.  The caller must __syncthreads() after it has finished reading a
  // chunk and before the next acquire (the chunk after next lands in the same buffer).
  __device__ Chunk acquire() {
    Chunk c;
    const LayerW l = L[cl];
    const int rc = min(l.rpc, l.O - cr);
    c.layer = cl; c.r0 = cr; c.r1 = cr + rc;
    c.w = cbuf ? buf1 : buf0;
    issue();
    if (inflight == 2) cp_async_wait<1>(); else cp_async_wait<0>();
    --inflight;
    __syncthreads();
    cbuf ^= 1;
    cr += rc;
    if (cr >= l.O) { ++cl; cr = 0; }
    return c;
  }
};

// y[r][o] = b[o] + sum_i x[r][i] * W[o][i] for the rows o of one chunk; a warp owns four outputs at a
// time, lanes over i (conflict-free shared-memory reads), shuffle reduction.
template <int R>
__device__ __forceinline__ void linear_fwd_chunk(const Chunk& c, int I, const float* xs, int ldx,
                                                 const float* __restrict__ b, float* ys, int ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int nrows = c.r1 - c.r0;
  for (int o0 = warp * 4; o0 < nrows; o0 += nw * 4) {
    const int no = min(4, nrows - o0);
    const float* w = c.w + (size_t)o0 * I;
    float acc[4][R];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < R; ++r) acc[q][r] = 0.f;
    // bias of the output this lane will write (lanes 0, 8, 16, 24 <-> outputs 0..3), requested before the dot
    // products so that its L2 latency is hidden
    const int qw = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    const bool writer = (lane & 7) == 0 && qw < no;
    const float bias = writer ? __ldg(b + c.r0 + o0 + qw) : 0.f;
    for (int i = lane; i < I; i += 32) {
      const float w0 = w[i];
      const float w1 = (no > 1) ? w[I + i] : 0.f;
      const float w2 = (no > 2) ? w[2 * I + i] : 0.f;
      const float w3 = (no > 3) ? w[3 * I + i] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float x = xs[r * ldx + i];
        acc[0][r] = fmaf(w0, x, acc[0][r]);
        acc[1][r] = fmaf(w1, x, acc[1][r]);
        acc[2][r] = fmaf(w2, x, acc[2][r]);
        acc[3][r] = fmaf(w3, x, acc[3][r]);
      }
    }
    // transposing butterfly: 6 shuffles reduce the four sums at once (instead of 4 x 5)
    const bool hi16 = lane & 16, hi8 = lane & 8;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float k0 = hi16 ? acc[2][r] : acc[0][r], k1 = hi16 ? acc[3][r] : acc[1][r];
      float t0 = hi16 ? acc[0][r] : acc[2][r], t1 = hi16 ? acc[1][r] : acc[3][r];
      k0 += __shfl_xor_sync(0xffffffffu, t0, 16);
      k1 += __shfl_xor_sync(0xffffffffu, t1, 16);
      float u = hi8 ? k1 : k0, v = hi8 ? k0 : k1;
      u += __shfl_xor_sync(0xffffffffu, v, 8);
      u += __shfl_xor_sync(0xffffffffu, u, 4);
      u += __shfl_xor_sync(0xffffffffu, u, 2);
      u += __shfl_xor_sync(0xffffffffu, u, 1);
      if (writer) ys[r * ldy + c.r0 + o0 + qw] = u + bias;
    }
  }
}
// y = x . W^T + b for layer `layer` of the stager's sequence (all its chunks).  Ends with a barrier.
template <int R>
__device__ __forceinline__ void layer_fwd(Stager& sg, int layer, const float* xs, int ldx, const float* __restrict__ b,
                                          float* ys, int ldy) {
  const int I = sg.L[layer].I;
  while (sg.cl == layer) {
    const Chunk c = sg.acquire();
    linear_fwd_chunk<R>(c, I, xs, ldx, b, ys, ldy);
    __syncthreads();
  }
}
// dx[r][i] = sum_o dy[r][o] * W[o][i] for layer `layer`: thread (i, part) walks the chunk's rows o == part
// (mod parts); the partial sums meet in shared memory in a FIXED order (chunk by chunk, part by part), so the
// result is bit-reproducible -- the item shards of a multi-GPU run compute the replicated small layers redundantly
// and must not drift apart.  dxs is zeroed first.  Ends with a barrier.
template <int R>
__device__ __forceinline__ void layer_bwd(Stager& sg, int layer, const float* dys, int ldy, float* dxs, int ldx) {
  const int I = sg.L[layer].I;
  for (int q = threadIdx.x; q < R * I; q += blockDim.x) dxs[(q / I) * ldx + (q % I)] = 0.f;
  const int lanes = min((int)blockDim.x, (I + 31) & ~31);   // threads over i (whole warps)
  const int parts = max(1, (int)blockDim.x / lanes);
  const int part = threadIdx.x / lanes, il = threadIdx.x - part * lanes;
  while (sg.cl == layer) {
    const Chunk c = sg.acquire();          // barrier inside: the zeroing above is visible
    const int nrows = c.r1 - c.r0;
    // parts > 1 implies lanes >= I: at most one i per thread, its partial sums stay in registers until its turn
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    if (part < parts) {
      for (int i = il; i < I; i += lanes) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        int o = part;
        for (; o + 3 * parts < nrows; o += 4 * parts) {
          const float w0 = c.w[(size_t)o * I + i], w1 = c.w[(size_t)(o + parts) * I + i];
          const float w2 = c.w[(size_t)(o + 2 * parts) * I + i], w3 = c.w[(size_t)(o + 3 * parts) * I + i];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float* d = dys + r * ldy + c.r0 + o;
            acc[r] = fmaf(w0, d[0], acc[r]);
            acc[r] = fmaf(w1, d[parts], acc[r]);
            acc[r] = fmaf(w2, d[2 * parts], acc[r]);
            acc[r] = fmaf(w3, d[3 * parts], acc[r]);
          }
        }
        for (; o < nrows; o += parts) {
          const float w0 = c.w[(size_t)o * I + i];
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(w0, dys[r * ldy + c.r0 + o], acc[r]);
        }
        if (parts == 1) {                  // single owner of column i: plain accumulation
#pragma unroll
          for (int r = 0; r < R; ++r) dxs[r * ldx + i] += acc[r];
        }
      }
    }
    if (parts > 1) {
      for (int p = 0; p < parts; ++p) {
        if (part == p && il < I) {
#pragma unroll
          for (int r = 0; r < R; ++r) dxs[r * ldx + il] += acc[r];
        }
        __syncthreads();
      }
    } else {
      __syncthreads();
    }
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

__device__ __forceinline__ float* align16f(float* p) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
}
// Row r of h1pre computed in place of a load: b1 + (1/len) * sum of the W1t rows of the set's items (the sparse
// first encoder layer, aae.py:132-135).  Warp w takes the items w, w + nw, ...; lanes over the hidden units;
// the per-warp partial sums meet in `scratch` ([nw][H] floats).  Ends with a barrier.
template <int R>
__device__ __forceinline__ void gather_rows(float* xs, int ld, const aae_bag& bag, const float* __restrict__ b1, int H,
                                            int row0, int B, float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = 0; r < R; ++r) {
    const int row = row0 + r;
    int s = 0, e = 0;
    if (row < B) { s = __ldg(bag.indptr + row); e = __ldg(bag.indptr + row + 1); }
    float* part = scratch + warp * H;
    if ((H & 3) == 0) {
      const int H4 = H >> 2;
      for (int c = lane; c < H4; c += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = s + warp; j < e; j += nw) {
          const int i = __ldg(bag.indices + j);
          if (i >= bag.v_begin && i < bag.v_end) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(bag.W1t + (size_t)(i - bag.v_begin) * H) + c);
            acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
          }
        }
        reinterpret_cast<float4*>(part)[c] = acc;
      }
    } else {
      for (int c = lane; c < H; c += 32) {
        float acc = 0.f;
        for (int j = s + warp; j < e; j += nw) {
          const int i = __ldg(bag.indices + j);
          if (i >= bag.v_begin && i < bag.v_end) acc += __ldg(bag.W1t + (size_t)(i - bag.v_begin) * H + c);
        }
        part[c] = acc;
      }
    }
    __syncthreads();
    const float scale = bag.normalize ? 1.0f / fmaxf((float)(e - s), 1e-12f) : 1.0f;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      float a = 0.f;
      for (int w = 0; w < nw; ++w) a += scratch[w * H + c];
      xs[r * ld + c] = (row < B) ? fmaf(a, scale, __ldg(b1 + c)) : 0.f;
    }
    __syncthreads();
  }
}
// h1pre rows of this CTA: gathered from the bag, or loaded
template <int R>
__device__ __forceinline__ void input_rows(float* xs, int ld, const aae_bag& bag, const float* __restrict__ h1pre,
                                           const float* __restrict__ b1, int H, int row0, int B, float* scratch) {
  if (bag.indptr) {
    gather_rows<R>(xs, ld, bag, b1, H, row0, B, scratch);
  } else {
    load_rows<R>(xs, ld, h1pre, H, row0, B);
    __syncthreads();
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

// ---------------------------------------------------------------------------------------------
// weight gradients of the small layers: dW[o,i] = sum_r dY[r,o] * X[r,i], db[o] = sum_r dY[r,o].
// Up to 8 jobs per launch; one thread per output element (i fastest -> coalesced X reads).
// ---------------------------------------------------------------------------------------------
struct WJob {
  const float* dY; int ldy;   // row pitch
  const float* X;  int ldx;   // X == nullptr -> bias job (X == 1)
  int rows, O, I;
  float* out;                 // [O, I] gradient (may be nullptr when Adam is fused)
  float *p, *m, *v;           // p != nullptr: Adam applied in place right after the reduction
  int which;                  // 0: gen_lr step size (enc_optim / dec_optim), 1: reg_lr (gen_optim / disc_optim)
  int begin;                  // first linear output index of this job
};
struct WJobs {
  WJob j[10];
  int n, total;
  const aae_step_state* st;
  int trace_id;
};
constexpr int WG_OUT = 64;    // outputs per CTA
constexpr int WG_PARTS = 4;   // threads per output (split of the batch rows)
__global__ void __launch_bounds__(WG_OUT * WG_PARTS) small_wgrad_kernel(WJobs jobs) {
  __shared__ float part_s[WG_PARTS][WG_OUT];
  const int el = threadIdx.x % WG_OUT, part = threadIdx.x / WG_OUT;
  const int idx = blockIdx.x * WG_OUT + el;
  trace_mark(jobs.trace_id, 0);
  const bool live = idx < jobs.total;
  int k = 0;
#pragma unroll
  for (int q = 1; q < 10; ++q)
    if (q < jobs.n && idx >= jobs.j[q].begin) k = q;
  const WJob& J = jobs.j[k];
  const int e = idx - J.begin;
  const int o = e / J.I, i = e - o * J.I;
  float acc = 0.f;
  if (live) {
    // rows part, part + 4, ...: 8 independent loads in flight per thread
    const float* dy = J.dY + o;
    int r = part;
    if (J.X) {
      const float* x = J.X + i;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (; r + 3 * WG_PARTS < J.rows; r += 4 * WG_PARTS) {
        a0 = fmaf(dy[(size_t)r * J.ldy], x[(size_t)r * J.ldx], a0);
        a1 = fmaf(dy[(size_t)(r + WG_PARTS) * J.ldy], x[(size_t)(r + WG_PARTS) * J.ldx], a1);
        a2 = fmaf(dy[(size_t)(r + 2 * WG_PARTS) * J.ldy], x[(size_t)(r + 2 * WG_PARTS) * J.ldx], a2);
        a3 = fmaf(dy[(size_t)(r + 3 * WG_PARTS) * J.ldy], x[(size_t)(r + 3 * WG_PARTS) * J.ldx], a3);
      }
      for (; r < J.rows; r += WG_PARTS) a0 = fmaf(dy[(size_t)r * J.ldy], x[(size_t)r * J.ldx], a0);
      acc = (a0 + a1) + (a2 + a3);
    } else {
      for (; r < J.rows; r += WG_PARTS) acc += dy[(size_t)r * J.ldy];
    }
  }
  part_s[part][el] = acc;
  __syncthreads();
  if (part == 0 && live) {
    float g = (part_s[0][el] + part_s[1][el]) + (part_s[2][el] + part_s[3][el]);
    if (J.out) J.out[e] = g;
    if (J.p) {
      AdamK ak = adam_load(jobs.st, J.which);
      float pp = J.p[e], mm = J.m[e], vv = J.v[e];
      adam_update(ak, g, pp, mm, vv);
      J.p[e] = pp; J.m[e] = mm; J.v[e] = vv;
    }
  }
  trace_mark(jobs.trace_id, 1);
}

// Adam target of a packed parameter block (nullptr p: gradient only)
struct OptBlock {
  float *p, *m, *v;
  int which;
};
static OptBlock opt_of(const aae_adam_block& a) { return OptBlock{a.p, a.m, a.v, a.which}; }

static void add_job(WJobs& js, const float* dY, int ldy, const float* X, int ldx, int rows, int O, int I, float* out,
                    const OptBlock& ob, size_t off) {
  WJob& j = js.j[js.n++];
  j.dY = dY; j.ldy = ldy; j.X = X; j.ldx = ldx; j.rows = rows; j.O = O; j.I = I;
  j.out = out ? out + off : nullptr;
  j.p = ob.p ? ob.p + off : nullptr;
  j.m = ob.p ? ob.m + off : nullptr;
  j.v = ob.p ? ob.v + off : nullptr;
  j.which = ob.which;
  j.begin = js.total;
  js.total += O * I;
}
static int launch_jobs(const WJobs& js, cudaStream_t s) {
  small_wgrad_kernel<<<cdiv(js.total, WG_OUT), WG_OUT * WG_PARTS, 0, s>>>(js);
  return check_launch("small_wgrad");
}

static inline int rows_per_cta(int B) { return B > 2048 ? 4 : 1; }
#define SCRATCH_FLOATS(d) ((MLP_THREADS / 32) * (d).H + 8)

}  // namespace aae

using namespace aae;

#define LAUNCH_R(kernel, B, smem_floats_per_row, stream, ...)                                          \
  do {                                                                                                 \
    int R_ = rows_per_cta(B);                                                                          \
    size_t smem_ = sizeof(float) * ((size_t)(smem_floats_per_row) * R_ + 2 * STAGE_FLOATS + SCRATCH_FLOATS(d)) + 64; \
    if (R_ == 1) {                                                                                     \
      cudaFuncSetAttribute(kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);        \
      kernel<1><<<cdiv(B, 1), MLP_THREADS, smem_, as_stream(stream)>>>(__VA_ARGS__);                   \
    } else {                                                                                           \
      cudaFuncSetAttribute(kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_);        \
      kernel<4><<<cdiv(B, 4), MLP_THREADS, smem_, as_stream(stream)>>>(__VA_ARGS__);                   \
    }                                                                                                  \
  } while (0)

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
