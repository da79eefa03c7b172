// K1/K2: the encoder's first layer on the sparse multi-hot input as a CSR embedding-bag, its
// weight gradient, and the Adam updates of W1t (touched rows sparse, untouched rows swept).
// Reference: aaerec/aae.py:132-135 (F.normalize(p=1) + lin1), :703/:741 (backward), :706/:741
// (enc_optim / gen_optim steps over the same parameters).
#include <stdlib.h>
#include <algorithm>
#include <stddef.h>
#include "common.cuh"

namespace aae {
AAE_DEFINE_TRACE_SETTER(trace_set_bag)

// ---------------------------------------------------------------------------------------------
// step state
// ---------------------------------------------------------------------------------------------
__global__ void step_state_init_kernel(aae_step_state* st, float gen_lr, float reg_lr, uint64_t seed) {
  st->t = 0;
  st->rng_step = 0;
  st->beta1 = 0.9f;
  st->beta2 = 0.999f;
  st->eps = 1e-8f;
  st->gen_lr = gen_lr;
  st->reg_lr = reg_lr;
  st->step_size_gen = gen_lr;
  st->step_size_reg = reg_lr;
  st->bc2_sqrt = 1.0f;
  st->seed = seed;
}
__global__ void step_tick_kernel(aae_step_state* st) {
  int t = st->t + 1;
  st->t = t;
  st->rng_step += 1;
  // torch: bias_correction1 = 1 - beta1 ** step (python doubles); step_size = lr / bias_correction1
  double bc1 = 1.0 - pow(0.9, (double)t);
  double bc2 = 1.0 - pow(0.999, (double)t);
  st->step_size_gen = (float)((double)st->gen_lr / bc1);
  st->step_size_reg = (float)((double)st->reg_lr / bc1);
  st->bc2_sqrt = (float)sqrt(bc2);
}


// ---------------------------------------------------------------------------------------------
// K1: one warp per set, lanes stride over the float4 columns of each gathered row.
// H % 4 == 0 fast path (400-byte rows, 16-byte aligned); scalar path otherwise.
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) bag_fwd_kernel(const int32_t* __restrict__ indptr,
                                                      const int32_t* __restrict__ indices, int B,
                                                      const float* __restrict__ W1t,
                                                      const float* __restrict__ b1, int H, int normalize,
                                                      int v_begin, int v_end, int add_bias,
                                                      float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  trace_mark(TR_BAG_FWD, 0);
  if (warp >= B) return;
  int s = indptr[warp], e = indptr[warp + 1];
  float scale = 1.0f;
  if (normalize) {
    // F.normalize(x, p=1): x / max(sum|x|, 1e-12); binary rows -> 1/len, empty row stays zero
    float len = (float)(e - s);
    scale = 1.0f / fmaxf(len, 1e-12f);
  }
  if (VEC4) {
    int H4 = H >> 2;
    for (int c0 = 0; c0 < H4; c0 += 32) {
      int c = c0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < H4) {
        int j = s;
        // 4 independent row loads in flight per lane
        for (; j + 4 <= e; j += 4) {
          int i0 = indices[j], i1 = indices[j + 1], i2 = indices[j + 2], i3 = indices[j + 3];
          float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
          if (i0 >= v_begin && i0 < v_end) r0 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i0 - v_begin) * H) + c);
          if (i1 >= v_begin && i1 < v_end) r1 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i1 - v_begin) * H) + c);
          if (i2 >= v_begin && i2 < v_end) r2 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i2 - v_begin) * H) + c);
          if (i3 >= v_begin && i3 < v_end) r3 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i3 - v_begin) * H) + c);
          acc.x += (r0.x + r1.x) + (r2.x + r3.x);
          acc.y += (r0.y + r1.y) + (r2.y + r3.y);
          acc.z += (r0.z + r1.z) + (r2.z + r3.z);
          acc.w += (r0.w + r1.w) + (r2.w + r3.w);
        }
        for (; j < e; ++j) {
          int i0 = indices[j];
          if (i0 >= v_begin && i0 < v_end) {
            float4 r0 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i0 - v_begin) * H) + c);
            acc.x += r0.x; acc.y += r0.y; acc.z += r0.z; acc.w += r0.w;
          }
        }
        float4 bb = make_float4(0, 0, 0, 0);
        if (add_bias) bb = __ldg(reinterpret_cast<const float4*>(b1) + c);
        float4 o = make_float4(fmaf(acc.x, scale, bb.x), fmaf(acc.y, scale, bb.y), fmaf(acc.z, scale, bb.z),
                               fmaf(acc.w, scale, bb.w));
        reinterpret_cast<float4*>(out + (size_t)warp * H)[c] = o;
      }
    }
  } else {
    for (int c = lane; c < H; c += 32) {
      float acc = 0.f;
      for (int j = s; j < e; ++j) {
        int i0 = indices[j];
        if (i0 >= v_begin && i0 < v_end) acc += __ldg(W1t + (size_t)(i0 - v_begin) * H + c);
      }
      out[(size_t)warp * H + c] = fmaf(acc, scale, add_bias ? b1[c] : 0.f);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Fused-step bookkeeping: ONE single-CTA launch builds the touched-row slots and the transposed (item ->
// batch rows) view of the batch.  It runs on a side branch of the step's graph, under the decoder kernel.
// Phases (separated by block barriers): (1) slots: first occurrence of an item claims a slot; (2) per entry:
// position among the item's occurrences; (3) exclusive scan of the counts; (4) fill csc_row.
// ---------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 1024;
__device__ __forceinline__ int row_of_entry(const int32_t* __restrict__ indptr, int B, int e) {
  int lo = 0, hi = B;                       // largest b with indptr[b] <= e
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(indptr + mid) <= e) lo = mid; else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(PREP_THREADS) batch_prepare_kernel(
    const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int B, int v_begin, int v_end,
    int32_t* slot_of, int32_t* uniq, int32_t* n_uniq, int32_t* cnt, int32_t* pos, int32_t* csc_off, int32_t* csc_row,
    int cap) {
  __shared__ int s_n;
  __shared__ int s_warp[PREP_THREADS / 32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nnz = __ldg(indptr + B);
  trace_mark(TR_PREP, 0);
  if (tid == 0) { s_n = 0; s_carry = 0; }
  __syncthreads();
  for (int e = tid; e < nnz; e += PREP_THREADS) {
    int i = __ldg(indices + e);
    if (i < v_begin || i >= v_end) continue;
    i -= v_begin;
    if (atomicCAS(&slot_of[i], -1, -2) == -1) {
      int s = atomicAdd(&s_n, 1);
      if (s < cap) { uniq[s] = i; cnt[s] = 0; }
      __stcg(&slot_of[i], s);
    }
  }
  __syncthreads();
  const int n = min(s_n, cap);
  for (int e = tid; e < nnz; e += PREP_THREADS) {
    int i = __ldg(indices + e);
    if (i < v_begin || i >= v_end) continue;
    int s = __ldcg(&slot_of[i - v_begin]);
    pos[e] = (s >= 0 && s < cap) ? atomicAdd(&cnt[s], 1) : -1;
  }
  __syncthreads();
  // exclusive scan of cnt[0..n) in chunks of PREP_THREADS
  for (int base = 0; base < n; base += PREP_THREADS) {
    const int s = base + tid;
    const int c = (s < n) ? __ldcg(&cnt[s]) : 0;
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry;
    const int incl = x + (warp ? s_warp[warp - 1] : 0) + carry;
    if (s < n) csc_off[s] = incl - c;
    __syncthreads();
    if (tid == PREP_THREADS - 1) s_carry = incl;
    __syncthreads();
  }
  if (tid == 0) { csc_off[n] = s_carry; *n_uniq = n; }
  __syncthreads();
  for (int e = tid; e < nnz; e += PREP_THREADS) {
    int i = __ldg(indices + e);
    if (i < v_begin || i >= v_end) continue;
    const int p = pos[e];
    if (p < 0) continue;
    const int s = __ldcg(&slot_of[i - v_begin]);
    csc_row[__ldcg(&csc_off[s]) + p] = row_of_entry(indptr, B, e);
  }
  trace_mark(TR_PREP, 1);
}

// K2 fused: one warp per touched item: gradient row summed over the item's batch rows (ascending row order
// when the item occurs in at most 32 rows, so the result does not depend on the atomics' arrival order in
// batch_prepare), then Adam on W/m/v in registers.  No gradient buffer, no floating-point atomics.
template <bool VEC4>
__global__ void __launch_bounds__(256) w1_rows_update_kernel(const int32_t* __restrict__ uniq,
                                                             const int32_t* __restrict__ n_uniq, int cap,
                                                             const int32_t* __restrict__ csc_off,
                                                             const int32_t* __restrict__ csc_row,
                                                             const int32_t* __restrict__ indptr, int normalize,
                                                             const float* __restrict__ dh1, float* __restrict__ W,
                                                             float* __restrict__ m, float* __restrict__ v, int H,
                                                             const aae_step_state* __restrict__ st, int which,
                                                             int32_t* last) {
  const AdamK k = adam_load(st, which);
  const int n = min(*n_uniq, cap);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  trace_mark(which ? TR_ROWS2 : TR_ROWS1, 0);
  constexpr int CPL = VEC4 ? 1 : 4;          // VEC4: one float4 per lane (H <= 128); else up to 4 strided floats
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const size_t row = (size_t)uniq[s] * H;
    const int o0 = csc_off[s], cn = csc_off[s + 1] - o0;
    // old parameter / moments: requested first, consumed after the gradient is summed
    float4 p4 = make_float4(0, 0, 0, 0), m4 = p4, v4 = p4;
    float ps[CPL], ms[CPL], vs[CPL];
    if (VEC4) {
      if (lane * 4 < H) {
        p4 = *reinterpret_cast<const float4*>(W + row + lane * 4);
        m4 = *reinterpret_cast<const float4*>(m + row + lane * 4);
        v4 = *reinterpret_cast<const float4*>(v + row + lane * 4);
      }
    } else {
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        int c = lane + 32 * q;
        ps[q] = ms[q] = vs[q] = 0.f;
        if (c < H) { ps[q] = W[row + c]; ms[q] = m[row + c]; vs[q] = v[row + c]; }
      }
    }
    float4 g4 = make_float4(0, 0, 0, 0);
    float gs[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) gs[q] = 0.f;
    for (int c0 = 0; c0 < cn; c0 += 32) {
      const int nn = min(32, cn - c0);
      int r = (lane < nn) ? __ldg(csc_row + o0 + c0 + lane) : 0x7fffffff;
      float sc = 1.0f;
      if (normalize && lane < nn) sc = 1.0f / fmaxf((float)(__ldg(indptr + r + 1) - __ldg(indptr + r)), 1e-12f);
      // rank of this lane's row among the chunk's rows (rows are distinct)
      int rank = 0;
      for (int j = 0; j < nn; ++j) rank += (__shfl_sync(0xffffffffu, r, j) < r) ? 1 : 0;
      // rows in ascending order (deterministic sum), eight gradient rows in flight at a time: the hottest items of
      // a Zipf batch sit in most of its rows, and their warps are the tail of this kernel
      if (VEC4) {
        for (int k0 = 0; k0 < nn; k0 += 8) {
          float4 d[8];
          float w8[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int kx = k0 + u;
            const int src = __ffs(__ballot_sync(0xffffffffu, lane < nn && rank == kx)) - 1;   // -1: past the end
            const int b = __shfl_sync(0xffffffffu, r, src & 31);
            w8[u] = (src >= 0) ? __shfl_sync(0xffffffffu, sc, src & 31) : 0.f;
            d[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src >= 0 && lane * 4 < H) d[u] = __ldg(reinterpret_cast<const float4*>(dh1 + (size_t)b * H + lane * 4));
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (k0 + u < nn) {
              g4.x = fmaf(d[u].x, w8[u], g4.x); g4.y = fmaf(d[u].y, w8[u], g4.y);
              g4.z = fmaf(d[u].z, w8[u], g4.z); g4.w = fmaf(d[u].w, w8[u], g4.w);
            }
          }
        }
      } else {
#pragma unroll 4
        for (int kx = 0; kx < nn; ++kx) {
          const int src = __ffs(__ballot_sync(0xffffffffu, lane < nn && rank == kx)) - 1;
          const int b = __shfl_sync(0xffffffffu, r, src);
          const float w = __shfl_sync(0xffffffffu, sc, src);
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            int c = lane + 32 * q;
            if (c < H) gs[q] = fmaf(__ldg(dh1 + (size_t)b * H + c), w, gs[q]);
          }
        }
      }
    }
    if (VEC4) {
      if (lane * 4 < H) {
        adam_update(k, g4.x, p4.x, m4.x, v4.x);
        adam_update(k, g4.y, p4.y, m4.y, v4.y);
        adam_update(k, g4.z, p4.z, m4.z, v4.z);
        adam_update(k, g4.w, p4.w, m4.w, v4.w);
        *reinterpret_cast<float4*>(W + row + lane * 4) = p4;
        *reinterpret_cast<float4*>(m + row + lane * 4) = m4;
        *reinterpret_cast<float4*>(v + row + lane * 4) = v4;
      }
    } else {
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        int c = lane + 32 * q;
        if (c < H) {
          adam_update(k, gs[q], ps[q], ms[q], vs[q]);
          W[row + c] = ps[q]; m[row + c] = ms[q]; v[row + c] = vs[q];
        }
      }
    }
    if (last && lane == 0) last[uniq[s]] = st->t;
  }
  trace_mark(which ? TR_ROWS2 : TR_ROWS1, 1);
}
// any H: lanes stride over the hidden units, one pass per 32 columns (rows re-walked per pass)
__global__ void __launch_bounds__(256) w1_rows_update_wide_kernel(const int32_t* __restrict__ uniq,
                                                                  const int32_t* __restrict__ n_uniq, int cap,
                                                                  const int32_t* __restrict__ csc_off,
                                                                  const int32_t* __restrict__ csc_row,
                                                                  const int32_t* __restrict__ indptr, int normalize,
                                                                  const float* __restrict__ dh1, float* __restrict__ W,
                                                                  float* __restrict__ m, float* __restrict__ v, int H,
                                                                  const aae_step_state* __restrict__ st, int which,
                                                             int32_t* last) {
  const AdamK k = adam_load(st, which);
  const int n = min(*n_uniq, cap);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const size_t row = (size_t)uniq[s] * H;
    const int o0 = csc_off[s], cn = csc_off[s + 1] - o0;
    for (int c = lane; c < H; c += 32) {
      float g = 0.f;
      for (int j = 0; j < cn; ++j) {
        const int b = __ldg(csc_row + o0 + j);
        const float sc = normalize ? 1.0f / fmaxf((float)(__ldg(indptr + b + 1) - __ldg(indptr + b)), 1e-12f) : 1.0f;
        g = fmaf(__ldg(dh1 + (size_t)b * H + c), sc, g);
      }
      float pp = W[row + c], mm = m[row + c], vv = v[row + c];
      adam_update(k, g, pp, mm, vv);
      W[row + c] = pp; m[row + c] = mm; v[row + c] = vv;
    }
    if (last && lane == 0) last[uniq[s]] = st->t;
  }
}

// End of a fused step: slots released, losses finalised, accumulators cleared, step state advanced for the
// next step.
__global__ void __launch_bounds__(256) step_finish_kernel(int32_t* slot_of, const int32_t* __restrict__ uniq,
                                                          const int32_t* __restrict__ n_uniq, int cap, double* sums,
                                                          int n_sums, double n_total, int B, float* out,
                                                          aae_step_state* st, float* ktab) {
  trace_mark(TR_FINISH, 0);
  int n = min(*n_uniq, cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) slot_of[uniq[s]] = -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = (float)(sums[0] / n_total);
    out[1] = (float)(sums[1] / (double)B);
    out[2] = (float)(sums[2] / (double)B);
    for (int i = 0; i < n_sums; ++i) sums[i] = 0.0;
    int t = st->t + 1;
    st->t = t;
    st->rng_step += 1;
    double bc1 = 1.0 - pow(0.9, (double)t);
    double bc2 = 1.0 - pow(0.999, (double)t);
    st->step_size_gen = (float)((double)st->gen_lr / bc1);
    st->step_size_reg = (float)((double)st->reg_lr / bc1);
    st->bc2_sqrt = (float)sqrt(bc2);
    if (ktab)
      reinterpret_cast<float4*>(ktab)[t & (AAE_KTAB_SLOTS - 1)] =
          make_float4(st->step_size_gen, st->step_size_reg, 1.0f / st->bc2_sqrt, 0.f);
  }
  trace_mark(TR_FINISH, 1);
}

// ---------------------------------------------------------------------------------------------
// Dense-Adam-equivalent sweep of the rows that are not in the batch: both optimizer states in one
// pass (5 reads + 5 writes per parameter = 40 B), streaming float4, evict-first loads.
// A warp owns 32 consecutive float4; the row (and so the touched test) is per float4.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) w1_sweep_untouched_kernel(const int32_t* __restrict__ slot_of, int r_begin,
                                                                 int r_end, int H, float* __restrict__ W,
                                                                 float* __restrict__ m1, float* __restrict__ v1,
                                                                 float* __restrict__ m2, float* __restrict__ v2,
                                                                 const aae_step_state* __restrict__ st) {
  AdamK k1 = adam_load(st, 0), k2 = adam_load(st, 1);
  trace_mark(TR_SWEEP, 0);
  int H4 = H >> 2;
  size_t begin = (size_t)r_begin * H4, end = (size_t)r_end * H4;
  for (size_t q = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < end;
       q += (size_t)gridDim.x * blockDim.x) {
    int row = (int)(q / H4);
    if (slot_of[row] >= 0) continue;
    float4 p = __ldcs(reinterpret_cast<const float4*>(W) + q);
    float4 a = __ldcs(reinterpret_cast<const float4*>(m1) + q);
    float4 b = __ldcs(reinterpret_cast<const float4*>(v1) + q);
    float4 c = __ldcs(reinterpret_cast<const float4*>(m2) + q);
    float4 d = __ldcs(reinterpret_cast<const float4*>(v2) + q);
    adam_update_zero(k1, p.x, a.x, b.x); adam_update_zero(k2, p.x, c.x, d.x);
    adam_update_zero(k1, p.y, a.y, b.y); adam_update_zero(k2, p.y, c.y, d.y);
    adam_update_zero(k1, p.z, a.z, b.z); adam_update_zero(k2, p.z, c.z, d.z);
    adam_update_zero(k1, p.w, a.w, b.w); adam_update_zero(k2, p.w, c.w, d.w);
    __stcs(reinterpret_cast<float4*>(W) + q, p);
    __stcs(reinterpret_cast<float4*>(m1) + q, a);
    __stcs(reinterpret_cast<float4*>(v1) + q, b);
    __stcs(reinterpret_cast<float4*>(m2) + q, c);
    __stcs(reinterpret_cast<float4*>(v2) + q, d);
  }
  trace_mark(TR_SWEEP, 1);
}
__global__ void __launch_bounds__(256) w1_sweep_untouched_scalar_kernel(const int32_t* __restrict__ slot_of,
                                                                        int r_begin, int r_end, int H, float* W,
                                                                        float* m1, float* v1, float* m2, float* v2,
                                                                        const aae_step_state* __restrict__ st) {
  AdamK k1 = adam_load(st, 0), k2 = adam_load(st, 1);
  size_t begin = (size_t)r_begin * H, end = (size_t)r_end * H;
  for (size_t q = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < end;
       q += (size_t)gridDim.x * blockDim.x) {
    if (slot_of[(int)(q / H)] >= 0) continue;
    float p = W[q], a = m1[q], b = v1[q], c = m2[q], d = v2[q];
    adam_update_zero(k1, p, a, b);
    adam_update_zero(k2, p, c, d);
    W[q] = p; m1[q] = a; v1[q] = b; m2[q] = c; v2[q] = d;
  }
}



}  // namespace aae

using namespace aae;

extern "C" {

int aae_step_state_init(aae_step_state* st, float gen_lr, float reg_lr, uint64_t seed, void* stream) {
  AAE_REQUIRE(st, "null state");
  step_state_init_kernel<<<1, 1, 0, as_stream(stream)>>>(st, gen_lr, reg_lr, seed);
  return check_launch("step_state_init");
}
int aae_step_tick(aae_step_state* st, void* stream) {
  AAE_REQUIRE(st, "null state");
  step_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(st);
  return check_launch("step_tick");
}


int aae_bag_fwd(const int32_t* indptr, const int32_t* indices, int B, const float* W1t, const float* b1, int H,
                int normalize, int v_begin, int v_end, int add_bias, float* out, void* stream) {
  AAE_REQUIRE(indptr && indices && W1t && b1 && out, "null pointer");
  AAE_REQUIRE(B > 0 && H > 0, "bad size");
  int threads = 256, blocks = cdiv((int64_t)B * 32, threads);
  if ((H & 3) == 0)
    bag_fwd_kernel<true><<<blocks, threads, 0, as_stream(stream)>>>(indptr, indices, B, W1t, b1, H, normalize,
                                                                   v_begin, v_end, add_bias, out);
  else
    bag_fwd_kernel<false><<<blocks, threads, 0, as_stream(stream)>>>(indptr, indices, B, W1t, b1, H, normalize,
                                                                    v_begin, v_end, add_bias, out);
  return check_launch("bag_fwd");
}

int aae_batch_prepare(const int32_t* indptr, const int32_t* indices, int B, int v_begin, int v_end, int32_t* slot_of,
                      int32_t* uniq, int32_t* n_uniq, int32_t* cnt, int32_t* pos, int32_t* csc_off, int32_t* csc_row,
                      int cap, void* stream) {
  AAE_REQUIRE(indptr && indices && slot_of && uniq && n_uniq && cnt && pos && csc_off && csc_row, "null pointer");
  AAE_REQUIRE(B > 0 && cap > 0, "bad size");
  batch_prepare_kernel<<<1, PREP_THREADS, 0, as_stream(stream)>>>(indptr, indices, B, v_begin, v_end, slot_of, uniq,
                                                                  n_uniq, cnt, pos, csc_off, csc_row, cap);
  return check_launch("batch_prepare");
}
int aae_w1_rows_update(const int32_t* uniq, const int32_t* n_uniq, int cap, const int32_t* csc_off,
                       const int32_t* csc_row, const int32_t* indptr, int normalize, const float* dh1, float* W,
                       float* m, float* v, int H, const aae_step_state* st, int which, int32_t* last, void* stream) {
  AAE_REQUIRE(uniq && n_uniq && csc_off && csc_row && indptr && dh1 && W && m && v && st, "null pointer");
  int blocks = std::min(8 * sm_count(), std::max(1, cdiv((int64_t)cap * 32, 256)));
  if ((H & 3) == 0 && H <= 128)
    w1_rows_update_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(uniq, n_uniq, cap, csc_off, csc_row, indptr,
                                                                      normalize, dh1, W, m, v, H, st, which, last);
  else if (H <= 128)
    w1_rows_update_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(uniq, n_uniq, cap, csc_off, csc_row, indptr,
                                                                       normalize, dh1, W, m, v, H, st, which, last);
  else
    w1_rows_update_wide_kernel<<<blocks, 256, 0, as_stream(stream)>>>(uniq, n_uniq, cap, csc_off, csc_row, indptr,
                                                                     normalize, dh1, W, m, v, H, st, which, last);
  return check_launch("w1_rows_update");
}
int aae_step_finish(int32_t* slot_of, const int32_t* uniq, const int32_t* n_uniq, int cap, double* sums, int n_sums,
                    double n_total, int B, float* losses, aae_step_state* st, float* ktab, void* stream) {
  AAE_REQUIRE(slot_of && uniq && n_uniq && sums && losses && st && n_sums >= 3, "null pointer");
  step_finish_kernel<<<std::min(4 * sm_count(), std::max(1, cdiv(cap, 256))), 256, 0, as_stream(stream)>>>(
      slot_of, uniq, n_uniq, cap, sums, n_sums, n_total, B, losses, st, ktab);
  return check_launch("step_finish");
}
int aae_w1_sweep_untouched(const int32_t* slot_of, int r_begin, int r_end, int H, float* W, float* m1, float* v1,
                           float* m2, float* v2, const aae_step_state* st, void* stream) {
  AAE_REQUIRE(slot_of && W && m1 && v1 && m2 && v2 && st, "null pointer");
  if (r_end <= r_begin) return AAE_OK;
  int blocks = 8 * sm_count();
  if ((H & 3) == 0)
    w1_sweep_untouched_kernel<<<blocks, 256, 0, as_stream(stream)>>>(slot_of, r_begin, r_end, H, W, m1, v1, m2, v2, st);
  else
    w1_sweep_untouched_scalar_kernel<<<blocks, 256, 0, as_stream(stream)>>>(slot_of, r_begin, r_end, H, W, m1, v1, m2,
                                                                           v2, st);
  return check_launch("w1_sweep_untouched");
}
} // extern "C"
namespace aae {
// ---------------------------------------------------------------------------------------------
// Device-side epoch feed (aae.py:815-823 replaces sklearn.utils.shuffle + X_shuf[s:e].toarray()): the whole CSR
// matrix (and the condition matrix) stays in HBM, the host uploads one permutation per epoch, and every batch's
// packed CSR rows are built here from perm[row0 .. row0+B).
//   kernel 1 (one block): row lengths -> exclusive scan -> out_indptr[0..B]
//   kernel 2 (warp per row): column indices of the row, and its condition row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) batch_gather_scan_kernel(const int64_t* __restrict__ indptr_all,
                                                                 const int32_t* __restrict__ perm, int64_t row0, int B,
                                                                 int32_t* __restrict__ out_indptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int r = base + tid;
    int len = 0;
    if (r < B) {
      const int64_t src = perm ? (int64_t)perm[row0 + r] : row0 + r;
      len = (int)(indptr_all[src + 1] - indptr_all[src]);
    }
    int x = len;                         // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;                // inclusive totals of the warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + x - len;
    if (r < B) out_indptr[r] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) out_indptr[B] = carry_s;
}
__global__ void __launch_bounds__(256) batch_gather_copy_kernel(const int64_t* __restrict__ indptr_all,
                                                                const int32_t* __restrict__ indices_all,
                                                                const int32_t* __restrict__ perm, int64_t row0, int B,
                                                                const int32_t* __restrict__ out_indptr,
                                                                int32_t* __restrict__ out_indices,
                                                                const float* __restrict__ cond_all, int D,
                                                                float* __restrict__ out_cond) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    const int64_t src = perm ? (int64_t)perm[row0 + r] : row0 + r;
    const int64_t p0 = indptr_all[src];
    const int len = (int)(indptr_all[src + 1] - p0);
    const int o0 = out_indptr[r];
    for (int j = lane; j < len; j += 32) out_indices[o0 + j] = indices_all[p0 + j];
    if (cond_all)
      for (int j = lane; j < D; j += 32) out_cond[(size_t)r * D + j] = cond_all[(size_t)src * D + j];
  }
}

}  // namespace aae
extern "C" {
} // extern "C"
namespace aae {
// ---------------------------------------------------------------------------------------------
// Input corruption of the denoising autoencoder (dae.py:48-52 zeros_noise: mask = torch.rand(batch.size()) <
// noise_factor; batch[mask] = 0 -- in place, so the BCE target of dae.py:198-200 is the corrupted batch too): every
// entry of the batch's CSR rows is dropped with probability p.  noise != NULL: the reference's own [B,V] uniform draws
// (oracle-RNG mode); NULL: Philox keyed by (seed, step, row, item).  Two kernels: kept-entry counts + scan, then a
// per-row ballot compaction (column order is preserved).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool corrupt_keep(const float* __restrict__ noise, int V, const aae_step_state* st, int row,
                                             int item, float p) {
  if (noise) return !(noise[(size_t)row * V + item] < p);
  uint4 r = philox4x32(make_uint4((uint32_t)item, st->rng_step, (uint32_t)row, 0xd0e5u),
                       make_uint2((uint32_t)st->seed, (uint32_t)(st->seed >> 32)));
  return !(u01(r.x) < p);
}
__global__ void __launch_bounds__(1024) batch_corrupt_scan_kernel(const int32_t* __restrict__ in_indptr,
                                                                  const int32_t* __restrict__ in_indices, int B, int V,
                                                                  float p, const float* __restrict__ noise,
                                                                  const aae_step_state* __restrict__ st,
                                                                  int32_t* __restrict__ out_indptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int r = base + tid;
    int len = 0;
    if (r < B)
      for (int j = in_indptr[r]; j < in_indptr[r + 1]; ++j) len += corrupt_keep(noise, V, st, r, in_indices[j], p) ? 1 : 0;
    int x = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int carry = carry_s;
    if (r < B) out_indptr[r] = carry + (warp ? warp_tot[warp - 1] : 0) + x - len;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) out_indptr[B] = carry_s;
}
__global__ void __launch_bounds__(256) batch_corrupt_copy_kernel(const int32_t* __restrict__ in_indptr,
                                                                 const int32_t* __restrict__ in_indices, int B, int V,
                                                                 float p, const float* __restrict__ noise,
                                                                 const aae_step_state* __restrict__ st,
                                                                 const int32_t* __restrict__ out_indptr,
                                                                 int32_t* __restrict__ out_indices) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    const int p0 = in_indptr[r], p1 = in_indptr[r + 1];
    int o = out_indptr[r];
    for (int j0 = p0; j0 < p1; j0 += 32) {
      const int j = j0 + lane;
      const int item = (j < p1) ? in_indices[j] : 0;
      const bool keep = (j < p1) && corrupt_keep(noise, V, st, r, item, p);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) out_indices[o + __popc(m & ((1u << lane) - 1u))] = item;
      o += __popc(m);
    }
  }
}
}  // namespace aae
extern "C" {
int aae_batch_corrupt(const int32_t* in_indptr, const int32_t* in_indices, int B, int V, float p, const float* noise,
                      const aae_step_state* st, int32_t* out_indptr, int32_t* out_indices, void* stream) {
  AAE_REQUIRE(in_indptr && in_indices && st && out_indptr && out_indices, "null pointer");
  AAE_REQUIRE(B > 0 && V > 0 && p >= 0.f && p <= 1.f, "bad argument");
  AAE_REQUIRE(in_indices != out_indices && in_indptr != out_indptr, "in-place corruption is not supported");
  batch_corrupt_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(in_indptr, in_indices, B, V, p, noise, st, out_indptr);
  const int blocks = std::min(4 * sm_count(), std::max(1, cdiv(B, 8)));
  batch_corrupt_copy_kernel<<<blocks, 256, 0, as_stream(stream)>>>(in_indptr, in_indices, B, V, p, noise, st, out_indptr,
                                                                  out_indices);
  return check_launch("batch_corrupt");
}
int aae_batch_gather(const int64_t* indptr_all, const int32_t* indices_all, const int32_t* perm, int64_t row0, int B,
                     int32_t* out_indptr, int32_t* out_indices, const float* cond_all, int D, float* out_cond,
                     void* stream) {
  AAE_REQUIRE(indptr_all && indices_all && out_indptr && out_indices, "null pointer");
  AAE_REQUIRE(B > 0 && row0 >= 0 && D >= 0, "bad size");
  AAE_REQUIRE(!cond_all || (out_cond && D > 0), "condition rows without an output buffer");
  batch_gather_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(indptr_all, perm, row0, B, out_indptr);
  const int blocks = std::min(4 * sm_count(), std::max(1, cdiv(B, 8)));
  batch_gather_copy_kernel<<<blocks, 256, 0, as_stream(stream)>>>(indptr_all, indices_all, perm, row0, B, out_indptr,
                                                                out_indices, cond_all, D, out_cond);
  return check_launch("batch_gather");
}
int aae_upload_batch(const int32_t* indptr_host, const int32_t* indices_host, int B, int nnz, int32_t* indptr,
                     int32_t* indices, void* stream) {
  AAE_REQUIRE(indptr_host && indices_host && indptr && indices, "null pointer");
  cudaError_t e;
  const ptrdiff_t dh = indices_host - indptr_host, dd = indices - indptr;
  if (dh == dd && dh >= B + 1 && dh <= B + 1 + 4096) {
    // packed layout (indices right behind the indptr block, same offset on both sides): ONE copy -- the second
    // H2D transfer costs a PCIe round trip (~6 us) in front of every step, the few stale indptr slots nothing
    e = cudaMemcpyAsync(indptr, indptr_host, sizeof(int32_t) * (size_t)(dh + nnz), cudaMemcpyHostToDevice,
                        as_stream(stream));
  } else {
    e = cudaMemcpyAsync(indptr, indptr_host, sizeof(int32_t) * (size_t)(B + 1), cudaMemcpyHostToDevice,
                        as_stream(stream));
    if (e == cudaSuccess && nnz > 0)
      e = cudaMemcpyAsync(indices, indices_host, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice,
                          as_stream(stream));
  }
  if (e != cudaSuccess) {
    set_error("upload_batch: %s", cudaGetErrorString(e));
    return AAE_E_CUDA;
  }
  return AAE_OK;
}

// Same copy, with the pinned slot chosen ON THE DEVICE from the step counter: slot = (t + bias) & 1.  One captured
// graph then serves both slots (alternating between two instantiated graphs costs ~100 us per launch).
__global__ void __launch_bounds__(256) copy_words_sel_kernel(const uint32_t* __restrict__ src0,
                                                             const uint32_t* __restrict__ src1, uint32_t* dst0,
                                                             uint32_t* dst1, int64_t n,
                                                             const aae_step_state* __restrict__ st, int bias) {
  const int sel = (st->t + bias) & 1;
  const uint32_t* __restrict__ src = sel ? src1 : src0;
  uint32_t* __restrict__ dst = sel ? dst1 : dst0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
int aae_copy_words_sel(const void* src0, const void* src1, void* dst0, void* dst1, int64_t n_words,
                       const aae_step_state* st, int bias, void* stream) {
  AAE_REQUIRE(src0 && src1 && dst0 && dst1 && st && n_words >= 0, "bad argument");
  if (n_words == 0) return AAE_OK;
  int blocks = (int)std::min<int64_t>(2 * sm_count(), (n_words + 255) / 256);
  copy_words_sel_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint32_t*>(src0), reinterpret_cast<const uint32_t*>(src1),
      reinterpret_cast<uint32_t*>(dst0), reinterpret_cast<uint32_t*>(dst1), n_words, st, bias);
  return check_launch("copy_words_sel");
}


}  // extern "C"
