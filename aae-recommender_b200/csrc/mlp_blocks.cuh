// Building blocks of the row-local small-layer kernels (K4): weight staging, linear forward / backward on rows held in
// shared memory, dropout + ReLU, the in-kernel bag gather, parameter-block views and the weight-gradient jobs.
// Included by mlp.cu (AAE / AutoEncoder phases) and siblings.cu (DecodingRecommender, VAE).
#pragma once
#include <algorithm>
#include "common.cuh"

namespace aae {

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
static __global__ void __launch_bounds__(WG_OUT * WG_PARTS) small_wgrad_kernel(WJobs jobs) {
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

