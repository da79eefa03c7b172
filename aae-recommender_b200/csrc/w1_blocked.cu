// Time-blocked dense Adam for the encoder's first layer W1t (+ its two Adam states).
//
// torch's dense Adam moves every row of enc.lin1.weight twice per partial_fit (enc_optim aae.py:706,
// gen_optim aae.py:741) even when the row's gradient is zero (momentum decay): 40 bytes of HBM traffic per
// parameter per step for rows that are not in the batch.  The zero-gradient update of a row depends on
// nothing but the row itself and the step's bias corrections, so it can be applied LATE without changing a
// single bit, as long as it is applied before the row is read again:
//   * last[r] = the last step whose updates (real or zero-gradient) are applied to row r;
//   * ktab[t % AAE_KTAB_SLOTS] = the per-step constants of step t (step sizes, 1/sqrt(1-b2^t));
//   * every step, the rows of ONE of G contiguous groups (group t % G) are brought up to date by replaying
//     their pending steps in registers (<= G sequential zero-gradient updates per element, identical fp32
//     operations in identical order to the dense sweep) -- HBM traffic 40/G bytes per parameter per step;
//   * the rows of the batch are caught up to step t-1 before the encoder gathers them (aae_w1_catchup), get
//     their real updates from aae_w1_rows_update, which stamps last[r] = t;
//   * aae_w1_flush brings every row up to date (before predict / weight export).
// G = 1 is the plain dense sweep.  Reference semantics: torch/optim/adam.py::_single_tensor_adam.
#include "common.cuh"

namespace aae {
AAE_DEFINE_TRACE_SETTER(trace_set_w1b)

__device__ __forceinline__ void ktab_store(float* ktab, const aae_step_state* st) {
  float4 e = make_float4(st->step_size_gen, st->step_size_reg, 1.0f / st->bc2_sqrt, 0.f);
  reinterpret_cast<float4*>(ktab)[st->t & (AAE_KTAB_SLOTS - 1)] = e;
}
__global__ void ktab_write_kernel(const aae_step_state* st, float* ktab) { ktab_store(ktab, st); }

// `n` pending zero-gradient steps first+1 .. first+n of one float4 of a row, both optimizer states, enc_optim
// before gen_optim inside a step (the order of aae.py:706 / :741)
__device__ __forceinline__ void replay4(const float* __restrict__ ktab, const aae_step_state* __restrict__ st, int first,
                                        int n, float4& p, float4& a, float4& b, float4& c, float4& d) {
  AdamK k1, k2;
  k1.w1 = k2.w1 = (float)(1.0 - 0.9);
  k1.beta2 = k2.beta2 = st->beta2;
  k1.w2 = k2.w2 = (float)(1.0 - 0.999);
  k1.eps = k2.eps = st->eps;
  for (int j = 1; j <= n; ++j) {
    const float4 e = __ldg(reinterpret_cast<const float4*>(ktab) + ((first + j) & (AAE_KTAB_SLOTS - 1)));
    k1.step_size = e.x; k2.step_size = e.y;
    k1.inv_bc2_sqrt = k2.inv_bc2_sqrt = e.z;
    adam_update_zero(k1, p.x, a.x, b.x); adam_update_zero(k2, p.x, c.x, d.x);
    adam_update_zero(k1, p.y, a.y, b.y); adam_update_zero(k2, p.y, c.y, d.y);
    adam_update_zero(k1, p.z, a.z, b.z); adam_update_zero(k2, p.z, c.z, d.z);
    adam_update_zero(k1, p.w, a.w, b.w); adam_update_zero(k2, p.w, c.w, d.w);
  }
}
__device__ __forceinline__ void replay1(const float* __restrict__ ktab, const aae_step_state* __restrict__ st, int first,
                                        int n, float& p, float& a, float& b, float& c, float& d) {
  AdamK k1, k2;
  k1.w1 = k2.w1 = (float)(1.0 - 0.9);
  k1.beta2 = k2.beta2 = st->beta2;
  k1.w2 = k2.w2 = (float)(1.0 - 0.999);
  k1.eps = k2.eps = st->eps;
  for (int j = 1; j <= n; ++j) {
    const float4 e = __ldg(reinterpret_cast<const float4*>(ktab) + ((first + j) & (AAE_KTAB_SLOTS - 1)));
    k1.step_size = e.x; k2.step_size = e.y;
    k1.inv_bc2_sqrt = k2.inv_bc2_sqrt = e.z;
    adam_update_zero(k1, p, a, b);
    adam_update_zero(k2, p, c, d);
  }
}

// Cold elements.  A zero-gradient step computes  m' = m - 0.1 m,  v' = beta2 v,  W' = fma(-step, m' / denom, W)  with
// denom >= eps.  With |m| <= 2^-110, 1/denom <= 1e8 (eps >= 1e-8) and step <= 2^10 the increment of W is below 2^-73 in
// magnitude: for |W| >= 2^-40 (ulp >= 2^-63) the fma rounds back to W exactly, and for m == +0 the increment is -0 and
// W' == W for every W.  m decays by 0.9 per step, so an item's row turns cold ~650 steps after its last occurrence
// (it never reaches 0: round-to-nearest parks it on a denormal) and stays cold until it is in a batch again.
constexpr float kColdM = 7.7037198e-34f;     // 2^-110
constexpr float kColdWMin = 9.0949470e-13f;  // 2^-40
__device__ __forceinline__ bool cold4(const float4& m) {
  return fmaxf(fmaxf(fabsf(m.x), fabsf(m.y)), fmaxf(fabsf(m.z), fabsf(m.w))) <= kColdM;   // NaN -> false
}
__device__ __forceinline__ bool zero4(const float4& x) {      // all four are +0 (bit pattern 0; -0 does not qualify)
  return (__float_as_uint(x.x) | __float_as_uint(x.y) | __float_as_uint(x.z) | __float_as_uint(x.w)) == 0u;
}

// One warp brings row r from step `from` to step `to` (to - from pending steps), lanes over the float4 columns.
__device__ __forceinline__ void replay_row(size_t row_off, int H, int from, int to, float* __restrict__ W,
                                           float* __restrict__ m1, float* __restrict__ v1, float* __restrict__ m2,
                                           float* __restrict__ v2, const aae_step_state* __restrict__ st,
                                           const float* __restrict__ ktab, int lane) {
  if ((H & 3) == 0) {
    const int H4 = H >> 2;
    for (int c4 = lane; c4 < H4; c4 += 32) {
      const size_t q = (row_off >> 2) + c4;
      float4 p = __ldcs(reinterpret_cast<const float4*>(W) + q);
      float4 a = __ldcs(reinterpret_cast<const float4*>(m1) + q);
      float4 b = __ldcs(reinterpret_cast<const float4*>(v1) + q);
      float4 c = __ldcs(reinterpret_cast<const float4*>(m2) + q);
      float4 d = __ldcs(reinterpret_cast<const float4*>(v2) + q);
      replay4(ktab, st, from, to - from, p, a, b, c, d);
      reinterpret_cast<float4*>(W)[q] = p;
      __stcs(reinterpret_cast<float4*>(m1) + q, a);
      __stcs(reinterpret_cast<float4*>(v1) + q, b);
      __stcs(reinterpret_cast<float4*>(m2) + q, c);
      __stcs(reinterpret_cast<float4*>(v2) + q, d);
    }
  } else {
    for (int cc = lane; cc < H; cc += 32) {
      const size_t q = row_off + cc;
      float p = W[q], a = m1[q], b = v1[q], c = m2[q], d = v2[q];
      replay1(ktab, st, from, to - from, p, a, b, c, d);
      W[q] = p; m1[q] = a; v1[q] = b; m2[q] = c; v2[q] = d;
    }
  }
}

// Rows of the batch -> up to date through step t-1, before the encoder gathers them.  One warp per CSR entry;
// the first warp to claim an item in this step (claim[i] <- t) does the work, duplicates across rows skip.
__global__ void __launch_bounds__(256) w1_catchup_kernel(const int32_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ indices, int B, int v_begin,
                                                         int v_end, int32_t* claim, float* __restrict__ W,
                                                         float* __restrict__ m1, float* __restrict__ v1,
                                                         float* __restrict__ m2, float* __restrict__ v2, int32_t* last,
                                                         int H, const aae_step_state* __restrict__ st,
                                                         const float* __restrict__ ktab) {
  trace_mark(TR_CATCHUP, 0);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nnz = __ldg(indptr + B);
  const int t = st->t;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nnz; e += warps) {
    int i = __ldg(indices + e);
    if (i < v_begin || i >= v_end) continue;
    i -= v_begin;
    int from = 0, mine = 0;
    if (lane == 0) {
      mine = atomicExch(&claim[i], t) != t;
      if (mine) from = last[i];
    }
    mine = __shfl_sync(0xffffffffu, mine, 0);
    if (!mine) continue;
    from = __shfl_sync(0xffffffffu, from, 0);
    if (from >= t - 1) continue;
    replay_row((size_t)i * H, H, from, t - 1, W, m1, v1, m2, v2, st, ktab, lane);
    if (lane == 0) last[i] = t - 1;
  }
  trace_mark(TR_CATCHUP, 1);
}

// Group sweep (flush == 0: rows of group t % G that are not in the batch, up to step t) or flush (every row, up
// to step t-1).  One warp per row; slim CTAs so that it runs beside the latency-bound tail of the step.
template <int T>
__global__ void __launch_bounds__(T) w1_sweep_blocked_kernel(const int32_t* __restrict__ slot_of, int Vloc, int H,
                                                             float* __restrict__ W, float* __restrict__ m1,
                                                             float* __restrict__ v1, float* __restrict__ m2,
                                                             float* __restrict__ v2, int32_t* last,
                                                             const aae_step_state* __restrict__ st,
                                                             const float* __restrict__ ktab, int G, int flush) {
  trace_mark(TR_SWEEP, 0);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int t = st->t;
  int r0 = 0, r1 = Vloc, to = t - 1;
  if (!flush) {
    const int rpg = (Vloc + G - 1) / G;
    const int g = t % G;
    r0 = g * rpg;
    r1 = min(Vloc, r0 + rpg);
    to = t;
  }
  if ((H & 3) == 0) {
    // Flat walk over the float4 columns of the row range: with one warp per row only H/4 of the 32 lanes work (25 of 32
    // at n_hidden 100) and the replay is MUFU-bound.  A row's columns may then belong to two warps, so `last` is not
    // written here (the second warp would find the row up to date and skip its part): w1_last_kernel does it afterwards.
    const int H4 = H >> 2;
    const long long n4 = (long long)(r1 - r0) * H4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    AdamK k1, k2;
    k1.w1 = k2.w1 = (float)(1.0 - 0.9);
    k1.beta2 = k2.beta2 = st->beta2;
    k1.w2 = k2.w2 = (float)(1.0 - 0.999);
    k1.eps = k2.eps = st->eps;
    // cold4()'s bound needs 1/denom <= 1/eps <= 1e8 and step sizes <= 2^10; the pending steps are earlier than step t,
    // their step sizes (lr / (1 - 0.9^t)) at most 10x the current ones
    const bool cold_ok = st->eps >= 1e-8f && st->step_size_gen <= 1.0f && st->step_size_reg <= 1.0f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
      const int r = r0 + (int)(e / H4);
      if (!flush && __ldg(slot_of + r) >= 0) continue;
      const int from = last[r];
      const int n = to - from;
      if (n <= 0) continue;
      const size_t q = (size_t)r0 * H4 + (size_t)e;
      float4 a = __ldcs(reinterpret_cast<const float4*>(m1) + q);
      float4 b = __ldcs(reinterpret_cast<const float4*>(v1) + q);
      float4 c = __ldcs(reinterpret_cast<const float4*>(m2) + q);
      float4 d = __ldcs(reinterpret_cast<const float4*>(v2) + q);
      float4 p;
      if (cold_ok && cold4(a) && cold4(c)) {
        // Cold elements (long-tail items: never in a batch, or not for hundreds of steps).  See cold4(): W cannot
        // change, so the sqrt / reciprocal / W update of every pending step is skipped; the moments still decay
        // through the same fp32 operations in the same order.  Results are bit-identical to the full replay.
        if (zero4(a) && zero4(c) && zero4(b) && zero4(d)) continue;    // never touched: every update is the identity
        bool w_safe = zero4(a) && zero4(c);
        if (!w_safe) {
          p = __ldcs(reinterpret_cast<const float4*>(W) + q);
          w_safe = fminf(fminf(fabsf(p.x), fabsf(p.y)), fminf(fabsf(p.z), fabsf(p.w))) >= kColdWMin;
        }
        if (w_safe) {
          for (int j = 1; j <= n; ++j) {
            a.x = fmaf(k1.w1, -a.x, a.x); a.y = fmaf(k1.w1, -a.y, a.y); a.z = fmaf(k1.w1, -a.z, a.z); a.w = fmaf(k1.w1, -a.w, a.w);
            c.x = fmaf(k2.w1, -c.x, c.x); c.y = fmaf(k2.w1, -c.y, c.y); c.z = fmaf(k2.w1, -c.z, c.z); c.w = fmaf(k2.w1, -c.w, c.w);
            b.x *= k1.beta2; b.y *= k1.beta2; b.z *= k1.beta2; b.w *= k1.beta2;
            d.x *= k2.beta2; d.y *= k2.beta2; d.z *= k2.beta2; d.w *= k2.beta2;
          }
          __stcs(reinterpret_cast<float4*>(m1) + q, a);
          __stcs(reinterpret_cast<float4*>(v1) + q, b);
          __stcs(reinterpret_cast<float4*>(m2) + q, c);
          __stcs(reinterpret_cast<float4*>(v2) + q, d);
          continue;
        }
      } else {
        p = __ldcs(reinterpret_cast<const float4*>(W) + q);
      }
      for (int j = 1; j <= n; ++j) {       // the same operations in the same order as replay4
        const float4 kk = __ldg(reinterpret_cast<const float4*>(ktab) + ((from + j) & (AAE_KTAB_SLOTS - 1)));
        k1.step_size = kk.x; k2.step_size = kk.y;
        k1.inv_bc2_sqrt = k2.inv_bc2_sqrt = kk.z;
        adam_update_zero(k1, p.x, a.x, b.x); adam_update_zero(k2, p.x, c.x, d.x);
        adam_update_zero(k1, p.y, a.y, b.y); adam_update_zero(k2, p.y, c.y, d.y);
        adam_update_zero(k1, p.z, a.z, b.z); adam_update_zero(k2, p.z, c.z, d.z);
        adam_update_zero(k1, p.w, a.w, b.w); adam_update_zero(k2, p.w, c.w, d.w);
      }
      reinterpret_cast<float4*>(W)[q] = p;
      __stcs(reinterpret_cast<float4*>(m1) + q, a);
      __stcs(reinterpret_cast<float4*>(v1) + q, b);
      __stcs(reinterpret_cast<float4*>(m2) + q, c);
      __stcs(reinterpret_cast<float4*>(v2) + q, d);
    }
  } else {
    for (int r = r0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < r1; r += warps) {
      if (!flush && __ldg(slot_of + r) >= 0) continue;
      const int from = last[r];
      if (from >= to) continue;
      replay_row((size_t)r * H, H, from, to, W, m1, v1, m2, v2, st, ktab, lane);
      if (lane == 0) last[r] = to;
    }
  }
  trace_mark(TR_SWEEP, 1);
}
// rows swept by the flat walk above -> last[r] = to
__global__ void __launch_bounds__(256) w1_last_kernel(const int32_t* __restrict__ slot_of, int Vloc, int32_t* last,
                                                      const aae_step_state* __restrict__ st, int G, int flush) {
  const int t = st->t;
  int r0 = 0, r1 = Vloc, to = t - 1;
  if (!flush) {
    const int rpg = (Vloc + G - 1) / G;
    r0 = (t % G) * rpg;
    r1 = min(Vloc, r0 + rpg);
    to = t;
  }
  for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x) {
    if (!flush && __ldg(slot_of + r) >= 0) continue;
    if (last[r] < to) last[r] = to;
  }
}

}  // namespace aae

using namespace aae;

extern "C" {

int aae_ktab_write(const aae_step_state* st, float* ktab, void* stream) {
  AAE_REQUIRE(st && ktab, "null pointer");
  ktab_write_kernel<<<1, 1, 0, as_stream(stream)>>>(st, ktab);
  return check_launch("ktab_write");
}

int aae_w1_catchup(const int32_t* indptr, const int32_t* indices, int B, int v_begin, int v_end, int32_t* claim, float* W,
                   float* m1, float* v1, float* m2, float* v2, int32_t* last, int H, const aae_step_state* st,
                   const float* ktab, void* stream) {
  AAE_REQUIRE(indptr && indices && claim && W && m1 && v1 && m2 && v2 && last && st && ktab, "null pointer");
  AAE_REQUIRE(B > 0 && H > 0, "bad size");
  int blocks = std::min(8 * sm_count(), std::max(1, cdiv((int64_t)B * 16 * 32, 256)));
  w1_catchup_kernel<<<blocks, 256, 0, as_stream(stream)>>>(indptr, indices, B, v_begin, v_end, claim, W, m1, v1, m2, v2,
                                                          last, H, st, ktab);
  return check_launch("w1_catchup");
}

int aae_w1_sweep_blocked(const int32_t* slot_of, int Vloc, int H, float* W, float* m1, float* v1, float* m2, float* v2,
                         int32_t* last, const aae_step_state* st, const float* ktab, int G, int flush, int ctas_per_sm,
                         void* stream) {
  AAE_REQUIRE(W && m1 && v1 && m2 && v2 && last && st && ktab, "null pointer");
  AAE_REQUIRE(flush || slot_of, "slot_of missing");
  AAE_REQUIRE(G >= 1 && G < AAE_KTAB_SLOTS, "G outside [1, AAE_KTAB_SLOTS)");
  if (Vloc <= 0) return AAE_OK;
  if (ctas_per_sm > 0)
    w1_sweep_blocked_kernel<128><<<ctas_per_sm * sm_count(), 128, 0, as_stream(stream)>>>(slot_of, Vloc, H, W, m1, v1, m2,
                                                                                       v2, last, st, ktab, G, flush);
  else
    w1_sweep_blocked_kernel<256><<<8 * sm_count(), 256, 0, as_stream(stream)>>>(slot_of, Vloc, H, W, m1, v1, m2, v2,
                                                                             last, st, ktab, G, flush);
  if ((H & 3) == 0) {
    const int rows = flush ? Vloc : (Vloc + G - 1) / G;
    w1_last_kernel<<<std::max(1, std::min(2 * sm_count(), (rows + 255) / 256)), 256, 0, as_stream(stream)>>>(slot_of, Vloc, last,
                                                                                                            st, G, flush);
  }
  return check_launch("w1_sweep_blocked");
}

}  // extern "C"
