// Item-sharded exchange over NVLink peer memory: the one-shot all-reduce(sum) of the [B,H] partial sums that
// the item shards of one box exchange per phase (X.W1^T partials twice, dh2 + the loss partial).
//
// The messages are 40 KB (B=100) .. 4 MB (B=10k): latency is everything, bandwidth nothing.  Every rank owns one
// cudaMalloc'ed exchange buffer that all peers map through CUDA IPC (NVSwitch gives every GPU a direct path to
// every peer).  One kernel per exchange, no host round trip, capturable in the step's CUDA graph:
//   1. publish : block b copies its chunk of the local partial into slot[e][seq&1] of the OWN buffer;
//   2. signal  : st.release.sys of `seq` into flags[e][b][my rank] of every PEER's buffer;
//   3. wait    : ld.acquire.sys on the own flags[e][b][r] until every peer r has published chunk b (bounded spin);
//   4. reduce  : the chunk is summed over the ranks IN RANK ORDER from the peers' slots (ld.volatile over NVLink),
//                so every rank gets bit-identical sums (the replicated small layers must not drift apart);
//   5. the last block to finish advances seq[e].
// Slots are double-buffered by the parity of seq: a rank overwrites slot[p] in exchange s+2 only after it passed
// the wait of exchange s+1, i.e. after every peer started s+1 and therefore finished reading slot[p] of exchange s.
// Replaces what the reference does implicitly inside one dense GEMM on one device (aae.py:132-135, 176-177, 703).
#include <string.h>
#include <stddef.h>
#include <algorithm>
#include "common.cuh"

namespace aae {

constexpr int PX_MAX_BLOCKS = 32;
constexpr int PX_EXTRA = 4;   // doubles exchanged beside the floats (loss partial sums)

struct PxHeader {
  uint32_t flags[AAE_PEER_EXCHANGES][PX_MAX_BLOCKS][AAE_PEER_MAX_WORLD];   // written by the peers
  uint32_t seq[AAE_PEER_EXCHANGES];                                        // completed exchanges (local)
  uint32_t ticket[AAE_PEER_EXCHANGES];                                     // blocks done in the running exchange
  uint32_t err;                                                            // a wait timed out
  uint32_t pad[7];
};

__host__ __device__ inline size_t px_align(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t px_slot_bytes(int64_t n_max) { return px_align((size_t)n_max * 4 + PX_EXTRA * 8); }
__host__ __device__ inline size_t px_header_bytes() { return px_align(sizeof(PxHeader)); }
__device__ __forceinline__ char* px_slot(void* base, int e, int par, int64_t n_max) {
  return reinterpret_cast<char*>(base) + px_header_bytes() + (size_t)(e * 2 + par) * px_slot_bytes(n_max);
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_volatile_f(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_volatile_d(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(aae_peers P, int e, float* __restrict__ data, int n,
                                                             double* __restrict__ extra, int n_extra, int64_t n_max,
                                                             long long spin_cycles) {
  PxHeader* my = reinterpret_cast<PxHeader*>(P.base[P.rank]);
  const int tid = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(&my->seq[e]) + 1u;
  const int par = (int)(seq & 1u);
  float* mine = reinterpret_cast<float*>(px_slot(P.base[P.rank], e, par, n_max));
  double* mine_x = reinterpret_cast<double*>(reinterpret_cast<char*>(mine) + (size_t)n_max * 4);
  // chunk of this block: the same float range on every rank (multiples of 4), float4 accesses when this rank's
  // pointer and the message length allow it
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
  const int per4 = (((n + 3) >> 2) + nb - 1) / nb;
  const int f0 = min(n, b * per4 * 4), f1 = min(n, f0 + per4 * 4);
  const int u0 = vec ? (f0 >> 2) : f0, u1 = vec ? (f1 >> 2) : f1;
  // 1. publish
  if (vec) {
    for (int u = u0 + tid; u < u1; u += blockDim.x)
      reinterpret_cast<float4*>(mine)[u] = reinterpret_cast<const float4*>(data)[u];
  } else {
    for (int u = u0 + tid; u < u1; u += blockDim.x) mine[u] = data[u];
  }
  if (b == 0 && tid < n_extra) mine_x[tid] = extra[tid];
  __shared__ int timed_out;
  if (tid == 0) timed_out = 0;
  __threadfence_system();
  __syncthreads();
  // 2. signal, 3. wait (one thread per peer)
  if (tid < P.world && tid != P.rank) {
    PxHeader* peer = reinterpret_cast<PxHeader*>(P.base[tid]);
    st_release_sys(&peer->flags[e][b][P.rank], seq);
    const uint32_t* f = &my->flags[e][b][tid];
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      if (clock64() - t0 > spin_cycles) {   // a peer died or diverged: report, do not hang the box
        my->err = 1u;
        timed_out = 1;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  // A timed-out wait is fatal for the step: the peers' slots may be stale, so the result is poisoned with NaN
  // (losses and weights turn NaN at once) instead of being summed from incomplete data; the host raises at its
  // next synchronisation point (AAEEngine.check_exchange).
  const float poison = timed_out ? __int_as_float(0x7fc00000) : 0.f;
  // 4. reduce in rank order
  if (vec) {
    for (int u = u0 + tid; u < u1; u += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < P.world; ++r) {
        const float4* src = reinterpret_cast<const float4*>(px_slot(P.base[r], e, par, n_max)) + u;
        const float4 x = (r == P.rank) ? *src : ld_volatile_f4(src);
        if (r == 0) acc = x;
        else { acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
      }
      if (timed_out) acc = make_float4(poison, poison, poison, poison);
      reinterpret_cast<float4*>(data)[u] = acc;
    }
  } else {
    for (int u = u0 + tid; u < u1; u += blockDim.x) {
      float acc = 0.f;
      for (int r = 0; r < P.world; ++r) {
        const float* src = reinterpret_cast<const float*>(px_slot(P.base[r], e, par, n_max)) + u;
        const float x = (r == P.rank) ? *src : ld_volatile_f(src);
        acc = (r == 0) ? x : acc + x;
      }
      data[u] = timed_out ? poison : acc;
    }
  }
  if (b == 0 && tid < n_extra) {
    double acc = 0.0;
    for (int r = 0; r < P.world; ++r) {
      const double* src =
          reinterpret_cast<const double*>(px_slot(P.base[r], e, par, n_max) + (size_t)n_max * 4) + tid;
      const double x = (r == P.rank) ? *src : ld_volatile_d(src);
      acc = (r == 0) ? x : acc + x;
    }
    extra[tid] = timed_out ? (double)poison : acc;
  }
  // 5. the last block to finish advances the sequence number (every block has read it by then)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(&my->ticket[e], 1u);
    if (done == (uint32_t)nb - 1u) {
      my->ticket[e] = 0u;
      *reinterpret_cast<volatile uint32_t*>(&my->seq[e]) = seq;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Encoder first layer of an item shard FUSED with its exchange (one launch instead of aae_bag_fwd + aae_peer_allreduce):
// block b gathers the partial sums of batch rows [8b, 8b+8) from the local W1t rows (one warp per row, as bag_fwd_kernel)
// straight into its exchange slot, signals, waits for the peers' block b, sums the rows over the ranks in rank order and
// adds the bias -- out[b,:] = b1 + sum_r partial_r[b,:], bit-identical on every rank.  n_hidden % 4 == 0, B <= 8 * 32.
// ---------------------------------------------------------------------------------------------
constexpr int PXB_ROWS = 8;
__global__ void __launch_bounds__(256) peer_bag_allreduce_kernel(aae_peers P, int e, const int32_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices, int B,
                                                                 const float* __restrict__ W1t,
                                                                 const float* __restrict__ b1, int H, int normalize,
                                                                 int v_begin, int v_end, float* __restrict__ out,
                                                                 int64_t n_max, long long spin_cycles) {
  PxHeader* my = reinterpret_cast<PxHeader*>(P.base[P.rank]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x, nb = gridDim.x;
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(&my->seq[e]) + 1u;
  const int par = (int)(seq & 1u);
  float* mine = reinterpret_cast<float*>(px_slot(P.base[P.rank], e, par, n_max));
  const int row = b * PXB_ROWS + warp;
  const int H4 = H >> 2;
  __shared__ int timed_out;
  if (tid == 0) timed_out = 0;
  // 1. gather + publish
  if (row < B) {
    const int s = indptr[row], en = indptr[row + 1];
    const float scale = normalize ? 1.0f / fmaxf((float)(en - s), 1e-12f) : 1.0f;
    for (int c = lane; c < H4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int j = s;
      for (; j + 4 <= en; j += 4) {
        const int i0 = indices[j], i1 = indices[j + 1], i2 = indices[j + 2], i3 = indices[j + 3];
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
      for (; j < en; ++j) {
        const int i0 = indices[j];
        if (i0 >= v_begin && i0 < v_end) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(W1t + (size_t)(i0 - v_begin) * H) + c);
          acc.x += r0.x; acc.y += r0.y; acc.z += r0.z; acc.w += r0.w;
        }
      }
      acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
      reinterpret_cast<float4*>(mine + (size_t)row * H)[c] = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal, 3. wait (one thread per peer)
  if (tid < P.world && tid != P.rank) {
    PxHeader* peer = reinterpret_cast<PxHeader*>(P.base[tid]);
    st_release_sys(&peer->flags[e][b][P.rank], seq);
    const uint32_t* f = &my->flags[e][b][tid];
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      if (clock64() - t0 > spin_cycles) {
        my->err = 1u;
        timed_out = 1;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  // 4. reduce the block's rows in rank order, add the bias
  if (row < B) {
    for (int c = lane; c < H4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < P.world; ++r) {
        const float4* src = reinterpret_cast<const float4*>(px_slot(P.base[r], e, par, n_max)) + (size_t)row * H4 + c;
        const float4 x = (r == P.rank) ? *src : ld_volatile_f4(src);
        if (r == 0) acc = x;
        else { acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
      }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b1) + c);
      acc.x += bb.x; acc.y += bb.y; acc.z += bb.z; acc.w += bb.w;
      if (timed_out) acc.x = acc.y = acc.z = acc.w = __int_as_float(0x7fc00000);
      reinterpret_cast<float4*>(out + (size_t)row * H)[c] = acc;
    }
  }
  // 5. the last block to finish advances the sequence number
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(&my->ticket[e], 1u);
    if (done == (uint32_t)nb - 1u) {
      my->ticket[e] = 0u;
      *reinterpret_cast<volatile uint32_t*>(&my->seq[e]) = seq;
    }
  }
}

}  // namespace aae

using namespace aae;

extern "C" {

int64_t aae_peer_buffer_bytes(int64_t n_max) {
  if (n_max <= 0) return 0;
  return (int64_t)(px_header_bytes() + (size_t)AAE_PEER_EXCHANGES * 2 * px_slot_bytes(n_max));
}

int aae_peer_alloc(int64_t n_max, void** base_out, unsigned char* handle_out) {
  AAE_REQUIRE(n_max > 0 && base_out && handle_out, "bad argument");
  const size_t bytes = (size_t)aae_peer_buffer_bytes(n_max);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("aae_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return AAE_E_CUDA; }
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("aae_peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    cudaGetLastError();
    return AAE_E_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == AAE_PEER_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *base_out = p;
  return AAE_OK;
}

int aae_peer_open(const unsigned char* handle, void** base_out) {
  AAE_REQUIRE(handle && base_out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("aae_peer_open: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return AAE_E_CUDA;
  }
  *base_out = p;
  return AAE_OK;
}

int aae_peer_close(void* base) {
  if (base) cudaIpcCloseMemHandle(base);
  cudaGetLastError();
  return AAE_OK;
}

int aae_peer_free(void* base) {
  if (base) cudaFree(base);
  cudaGetLastError();
  return AAE_OK;
}

int aae_peer_error(const void* base, int* err_host) {
  AAE_REQUIRE(base && err_host, "null pointer");
  uint32_t v = 0;
  cudaError_t e = cudaMemcpy(&v, reinterpret_cast<const char*>(base) + offsetof(PxHeader, err), 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error("aae_peer_error: %s", cudaGetErrorString(e)); return AAE_E_CUDA; }
  *err_host = (int)v;
  return AAE_OK;
}

int aae_peer_bag_allreduce(aae_peers peers, int exchange, const int32_t* indptr, const int32_t* indices, int B,
                           const float* W1t, const float* b1, int H, int normalize, int v_begin, int v_end, float* out,
                           int64_t n_max, void* stream) {
  AAE_REQUIRE(peers.world >= 2 && peers.world <= AAE_PEER_MAX_WORLD, "world outside [2, AAE_PEER_MAX_WORLD]");
  AAE_REQUIRE(peers.rank >= 0 && peers.rank < peers.world, "bad rank");
  AAE_REQUIRE(exchange >= 0 && exchange < AAE_PEER_EXCHANGES, "bad exchange id");
  AAE_REQUIRE(indptr && indices && W1t && b1 && out, "null pointer");
  AAE_REQUIRE(B > 0 && B <= PXB_ROWS * PX_MAX_BLOCKS && (H & 3) == 0 && (int64_t)B * H <= n_max,
              "fused gather + exchange handles batches of up to 256 rows, n_hidden % 4 == 0");
  for (int r = 0; r < peers.world; ++r) AAE_REQUIRE(peers.base[r], "peer buffer not mapped");
  const int blocks = cdiv(B, PXB_ROWS);
  peer_bag_allreduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(peers, exchange, indptr, indices, B, W1t, b1, H,
                                                                  normalize, v_begin, v_end, out, n_max, 40000000000LL);
  return check_launch("peer_bag_allreduce");
}

int aae_peer_allreduce(aae_peers peers, int exchange, float* data, int n, double* extra, int n_extra, int64_t n_max,
                       void* stream) {
  AAE_REQUIRE(peers.world >= 1 && peers.world <= AAE_PEER_MAX_WORLD, "world outside [1, AAE_PEER_MAX_WORLD]");
  AAE_REQUIRE(peers.rank >= 0 && peers.rank < peers.world, "bad rank");
  AAE_REQUIRE(exchange >= 0 && exchange < AAE_PEER_EXCHANGES, "bad exchange id");
  AAE_REQUIRE(data && n > 0 && n <= n_max, "bad message");
  AAE_REQUIRE(n_extra >= 0 && n_extra <= PX_EXTRA && (n_extra == 0 || extra), "bad extra");
  for (int r = 0; r < peers.world; ++r) AAE_REQUIRE(peers.base[r], "peer buffer not mapped");
  if (peers.world == 1) return AAE_OK;
  const int blocks = std::max(1, std::min(PX_MAX_BLOCKS, cdiv((n + 3) >> 2, 256)));
  // ~20 s at 1.9 GHz: beyond any legitimate skew between the ranks of one box (graph capture, workspace growth)
  peer_allreduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(peers, exchange, data, n, extra, n_extra, n_max,
                                                            40000000000LL);
  return check_launch("peer_allreduce");
}

}  // extern "C"
