// K1/K2: the encoder's first layer on the sparse multi-hot input as a CSR embedding-bag, its
// weight gradient, and the Adam updates of W1t (touched rows sparse, untouched rows swept).
// Reference: aaerec/aae.py:132-135 (F.normalize(p=1) + lin1), :703/:741 (backward), :706/:741
// (enc_optim / gen_optim steps over the same parameters).
#include "common.cuh"

namespace aae {

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

// Start of one partial_fit in a single launch: Adam step counter / bias corrections, loss accumulators,
// the touched-row counter, and the zero fill of the accumulation buffers (dh2 and the two compact
// first-layer gradient buffers, nnz*H floats each).
__global__ void __launch_bounds__(256) step_begin_kernel(aae_step_state* st, double* sums, int n_sums, int32_t* counter,
                                                         float* z0, int64_t n0, float* z1, float* z2,
                                                         const int32_t* __restrict__ indptr, int B, int H) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int t = st->t + 1;
    st->t = t;
    st->rng_step += 1;
    double bc1 = 1.0 - pow(0.9, (double)t);
    double bc2 = 1.0 - pow(0.999, (double)t);
    st->step_size_gen = (float)((double)st->gen_lr / bc1);
    st->step_size_reg = (float)((double)st->reg_lr / bc1);
    st->bc2_sqrt = (float)sqrt(bc2);
    for (int i = 0; i < n_sums; ++i) sums[i] = 0.0;
    if (counter) *counter = 0;
  }
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n0; i += nth) z0[i] = 0.f;
  if (z1 || z2) {
    const int64_t n = (int64_t)indptr[B] * H;
    for (int64_t i = tid; i < n; i += nth) {
      if (z1) z1[i] = 0.f;
      if (z2) z2[i] = 0.f;
    }
  }
}
// End of one partial_fit: release the touched-row slots and finalise the three losses.
__global__ void __launch_bounds__(256) step_end_kernel(int32_t* slot_of, const int32_t* __restrict__ uniq,
                                                       const int32_t* __restrict__ n_uniq, int cap, const double* sums,
                                                       double n_total, int B, float* out) {
  int n = min(*n_uniq, cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) slot_of[uniq[s]] = -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[0] = (float)(sums[0] / n_total);
    out[1] = (float)(sums[1] / (double)B);
    out[2] = (float)(sums[2] / (double)B);
  }
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
// touched-row slots
// ---------------------------------------------------------------------------------------------
__global__ void batch_slots_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int B,
                                   int v_begin, int v_end, int32_t* slot_of, int32_t* uniq, int32_t* n_uniq) {
  int nnz = indptr[B];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) {
    int i = indices[e];
    if (i < v_begin || i >= v_end) continue;
    i -= v_begin;
    if (atomicCAS(&slot_of[i], -1, -2) == -1) {
      int s = atomicAdd(n_uniq, 1);
      uniq[s] = i;
      slot_of[i] = s;  // nobody reads slot_of before the next kernel
    }
  }
}
__global__ void batch_slots_reset_kernel(int32_t* slot_of, const int32_t* __restrict__ uniq, int32_t* n_uniq,
                                         int cap) {
  int n = min(*n_uniq, cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) slot_of[uniq[s]] = -1;
}
__global__ void zero_counter_kernel(int32_t* n_uniq) { *n_uniq = 0; }

__global__ void zero_rows_kernel(float* G, const int32_t* __restrict__ indptr, int B, int H) {
  size_t n = (size_t)indptr[B] * H;
  size_t n4 = n >> 2;
  float4* G4 = reinterpret_cast<float4*>(G);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    G4[i] = make_float4(0, 0, 0, 0);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) G[(n4 << 2) + threadIdx.x] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// K2: scatter-add of the first layer's weight gradient into the compact touched-row buffer.
// One warp per (set, item) pair (the set of a CSR entry is found by bisection of indptr); lanes over the
// hidden units, red.global.add.f32 (rows of popular items are shared by several sets).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bag_bwd_kernel(const int32_t* __restrict__ indptr,
                                                      const int32_t* __restrict__ indices, int B,
                                                      const float* __restrict__ dh1, int H, int normalize,
                                                      const int32_t* __restrict__ slot_of, int v_begin, int v_end,
                                                      float* __restrict__ G) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nnz = indptr[B];
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nnz; e += warps) {
    const int i = indices[e];
    if (i < v_begin || i >= v_end) continue;
    int lo = 0, hi = B;                       // largest b with indptr[b] <= e
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (indptr[mid] <= e) lo = mid; else hi = mid;
    }
    const int b = lo;
    const float scale = normalize ? 1.0f / fmaxf((float)(indptr[b + 1] - indptr[b]), 1e-12f) : 1.0f;
    float* grow = G + (size_t)slot_of[i - v_begin] * H;
    const float* drow = dh1 + (size_t)b * H;
    for (int c = lane; c < H; c += 32) atomicAdd(grow + c, drow[c] * scale);
  }
}

__global__ void __launch_bounds__(256) rows_adam_kernel(const int32_t* __restrict__ uniq,
                                                        const int32_t* __restrict__ n_uniq, int cap,
                                                        const float* __restrict__ G, float* __restrict__ W,
                                                        float* __restrict__ m, float* __restrict__ v, int H,
                                                        const aae_step_state* __restrict__ st, int which) {
  AdamK k = adam_load(st, which);
  int n = min(*n_uniq, cap);
  int lane = threadIdx.x & 31;
  int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    size_t row = (size_t)uniq[s] * H;
    for (int c = lane; c < H; c += 32) {
      float p = W[row + c], mm = m[row + c], vv = v[row + c];
      adam_update(k, G[(size_t)s * H + c], p, mm, vv);
      W[row + c] = p;
      m[row + c] = mm;
      v[row + c] = vv;
    }
  }
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

__global__ void __launch_bounds__(256) adam_dense_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                         const aae_step_state* __restrict__ st, int which) {
  AdamK k = adam_load(st, which);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_update(k, g[i], pp, mm, vv);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

__global__ void finish_losses_kernel(const double* sums, double n_total, int B, float* out) {
  out[0] = (float)(sums[0] / n_total);
  out[1] = (float)(sums[1] / (double)B);
  out[2] = (float)(sums[2] / (double)B);
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

int aae_step_begin(aae_step_state* st, double* loss_sums, int n_sums, int32_t* n_uniq, float* dh2, int64_t n_dh2,
                   float* G1, float* G2, const int32_t* indptr, int B, int H, void* stream) {
  AAE_REQUIRE(st && loss_sums && n_sums >= 0, "null pointer");
  AAE_REQUIRE((!G1 && !G2) || indptr, "indptr missing");
  step_begin_kernel<<<2 * sm_count(), 256, 0, as_stream(stream)>>>(st, loss_sums, n_sums, n_uniq, dh2, dh2 ? n_dh2 : 0,
                                                                  G1, G2, indptr, B, H);
  return check_launch("step_begin");
}
int aae_step_end(int32_t* slot_of, const int32_t* uniq, const int32_t* n_uniq, int cap, const double* sums,
                 double n_total, int B, float* losses, void* stream) {
  AAE_REQUIRE(slot_of && uniq && n_uniq && sums && losses, "null pointer");
  step_end_kernel<<<std::min(4 * sm_count(), std::max(1, cdiv(cap, 256))), 256, 0, as_stream(stream)>>>(
      slot_of, uniq, n_uniq, cap, sums, n_total, B, losses);
  return check_launch("step_end");
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

int aae_batch_slots(const int32_t* indptr, const int32_t* indices, int B, int v_begin, int v_end, int32_t* slot_of,
                    int32_t* uniq, int32_t* n_uniq, int counter_is_zero, void* stream) {
  AAE_REQUIRE(indptr && indices && slot_of && uniq && n_uniq, "null pointer");
  if (!counter_is_zero) zero_counter_kernel<<<1, 1, 0, as_stream(stream)>>>(n_uniq);
  batch_slots_kernel<<<std::min(4 * sm_count(), std::max(1, cdiv((int64_t)B * 16, 256))), 256, 0, as_stream(stream)>>>(
      indptr, indices, B, v_begin, v_end, slot_of, uniq, n_uniq);
  return check_launch("batch_slots");
}
int aae_batch_slots_reset(int32_t* slot_of, const int32_t* uniq, int32_t* n_uniq, int cap, void* stream) {
  AAE_REQUIRE(slot_of && uniq && n_uniq, "null pointer");
  batch_slots_reset_kernel<<<std::min(4 * sm_count(), std::max(1, cdiv(cap, 256))), 256, 0, as_stream(stream)>>>(
      slot_of, uniq, n_uniq, cap);
  return check_launch("batch_slots_reset");
}
int aae_zero_rows(float* G, const int32_t* indptr, int B, int H, void* stream) {
  AAE_REQUIRE(G && indptr, "null pointer");
  zero_rows_kernel<<<2 * sm_count(), 256, 0, as_stream(stream)>>>(G, indptr, B, H);
  return check_launch("zero_rows");
}
int aae_bag_bwd(const int32_t* indptr, const int32_t* indices, int B, const float* dh1, int H, int normalize,
                const int32_t* slot_of, int v_begin, int v_end, float* G, void* stream) {
  AAE_REQUIRE(indptr && indices && dh1 && slot_of && G, "null pointer");
  int blocks = std::min(8 * sm_count(), std::max(1, cdiv((int64_t)B * 16 * 32, 256)));   // ~one warp per entry
  bag_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(indptr, indices, B, dh1, H, normalize, slot_of, v_begin,
                                                       v_end, G);
  return check_launch("bag_bwd");
}
int aae_rows_adam(const int32_t* uniq, const int32_t* n_uniq, int cap, const float* G, float* W, float* m, float* v,
                  int H, const aae_step_state* st, int which, void* stream) {
  AAE_REQUIRE(uniq && n_uniq && G && W && m && v && st, "null pointer");
  int blocks = std::min(8 * sm_count(), std::max(1, cdiv((int64_t)cap * 32, 256)));
  rows_adam_kernel<<<blocks, 256, 0, as_stream(stream)>>>(uniq, n_uniq, cap, G, W, m, v, H, st, which);
  return check_launch("rows_adam");
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
int aae_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, const aae_step_state* st, int which,
                   void* stream) {
  AAE_REQUIRE(p && g && m && v && st, "null pointer");
  int blocks = std::max(1, std::min(8 * sm_count(), cdiv(n, 256)));
  adam_dense_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, st, which);
  return check_launch("adam_dense");
}
int aae_upload_batch(const int32_t* indptr_host, const int32_t* indices_host, int B, int nnz, int32_t* indptr,
                     int32_t* indices, void* stream) {
  AAE_REQUIRE(indptr_host && indices_host && indptr && indices, "null pointer");
  cudaError_t e = cudaMemcpyAsync(indptr, indptr_host, sizeof(int32_t) * (size_t)(B + 1), cudaMemcpyHostToDevice,
                                  as_stream(stream));
  if (e == cudaSuccess && nnz > 0)
    e = cudaMemcpyAsync(indices, indices_host, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice,
                        as_stream(stream));
  if (e != cudaSuccess) {
    set_error("upload_batch: %s", cudaGetErrorString(e));
    return AAE_E_CUDA;
  }
  return AAE_OK;
}
int aae_finish_losses(const double* sums, double n_total, int B, float* out, void* stream) {
  AAE_REQUIRE(sums && out, "null pointer");
  finish_losses_kernel<<<1, 1, 0, as_stream(stream)>>>(sums, n_total, B, out);
  return check_launch("finish_losses");
}

}  // extern "C"
