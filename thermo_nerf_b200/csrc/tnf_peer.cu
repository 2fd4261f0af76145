// Multi-GPU exchange step of the training path: the gradient all-reduce that nerfstudio's DDP
// performs between backward and optimizer.step, fused with the Adam update of
// thermo_nerf/thermal_nerf/config_thermal_nerf.py:32-45 into ONE kernel over NVLink peer memory.
//
//   every rank owns 1/N of the flat parameter arena.  tnf_peer_adam_kernel on rank r
//     - reads shard r of EVERY rank's gradient arena (its own from HBM, the others as peer loads over
//       NVLink / NVSwitch) and averages them            -> reduce-scatter(mean)
//     - applies Adam to shard r (exp_avg / exp_avg_sq exist only for the owned shard)
//     - stores the updated shard into EVERY rank's parameter arena (peer stores)   -> all-gather
//   so the gradient bytes cross the fabric once in each direction, the optimizer state is sharded N ways
//   and no rank re-computes another rank's update.  Two flag barriers (system-scope release/acquire on
//   peer-mapped words) order it against the backward before and the forward after.
//
// The sum over ranks runs in rank order on the owner only, so parameters stay bit-identical on all ranks.
#include <cstring>

#include "tnf_device.cuh"
#include <cstdlib>

#include "tnf_host.h"

namespace tnf {

struct PeerAdamArgs {
  TnfPeerArena a;
  long long shard_begin, shard_end;  // float indices (multiples of 4) of the shard this rank owns
  float* m;                          // exp_avg of the shard    [shard_end - shard_begin]
  float* v;                          // exp_avg_sq of the shard
  long long seg_end[TNF_MAX_ADAM_SEGMENTS];
  float seg_step_size[TNF_MAX_ADAM_SEGMENTS];     // lr / bias_correction1
  float seg_inv_sqrt_bc2[TNF_MAX_ADAM_SEGMENTS];
  int seg_active[TNF_MAX_ADAM_SEGMENTS];
  int nseg;
  int push;  // 1: store the updated shard into every rank's parameters; 0: own copy only (peers pull it)
  float beta2, eps, omb1, omb2, inv_world;
};

__device__ __forceinline__ void peer_adam_one(float& p, float g, float& m, float& v, const PeerAdamArgs& a,
                                              const float step_size, const float inv_sqrt_bc2) {
  m = m + a.omb1 * (g - m);
  v = a.beta2 * v + a.omb2 * (g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) tnf_peer_adam_kernel(const __grid_constant__ PeerAdamArgs A) {
  const long long n4 = (A.shard_end - A.shard_begin) >> 2;
  const int W = A.a.world_size, me = A.a.rank;
  float4* __restrict__ M = reinterpret_cast<float4*>(A.m);
  float4* __restrict__ V = reinterpret_cast<float4*>(A.v);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e = A.shard_begin + 4 * i;
    int s = 0;
    while (s + 1 < A.nseg && e >= A.seg_end[s]) ++s;
    if (!A.seg_active[s]) continue;
    // reduce-scatter: this element of every rank's gradient arena (cache-volatile: peer lines are only
    // L1-cacheable here and must never be served stale)
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = 0; r < W; ++r) {
      const float4 t = __ldcv(reinterpret_cast<const float4*>(A.a.grads[r] + e));
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    g.x *= A.inv_world; g.y *= A.inv_world; g.z *= A.inv_world; g.w *= A.inv_world;
    float4 p = *reinterpret_cast<const float4*>(A.a.params[me] + e);
    float4 m = M[i], v = V[i];
    // never-touched entries (g = m = v = 0 on every rank): the update is exactly zero and every rank already holds
    // the same value - nothing to store, nothing to send
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f && m.w == 0.f &&
        v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f)
      continue;
    const float ss = A.seg_step_size[s], ib = A.seg_inv_sqrt_bc2[s];
    peer_adam_one(p.x, g.x, m.x, v.x, A, ss, ib);
    peer_adam_one(p.y, g.y, m.y, v.y, A, ss, ib);
    peer_adam_one(p.z, g.z, m.z, v.z, A, ss, ib);
    peer_adam_one(p.w, g.w, m.w, v.w, A, ss, ib);
    M[i] = m;
    V[i] = v;
    // all-gather, push flavour: the updated shard goes to every rank's parameter arena
    if (A.push) {
#pragma unroll 8
      for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(A.a.params[r] + e) = p;
    } else {
      *reinterpret_cast<float4*>(A.a.params[me] + e) = p;
    }
  }
}

// NVLS flavour: the same exchange through the NVSwitch multicast objects of the gradient and the parameter arenas.
//   reduce-scatter  multimem.ld_reduce.add: ONE load returns the sum of this element over all ranks, formed inside
//                   the switch - a rank receives 1/N of the arena instead of (N-1)/N of it
//   all-gather      multimem.st: ONE store is replicated by the switch into every rank's parameter arena - a rank
//                   sends 1/N of the arena instead of (N-1)/N of it
// Both directions of every link carry data at the same time (reduce traffic towards the owner, broadcast traffic
// away from it) and a chunk is stored as soon as it is updated, so the whole exchange is one pass over the shard.
// The switch's summation order differs from the rank-order sum of tnf_peer_adam_kernel (last-bit differences);
// every rank still receives the owner's one result, so the replicas stay bit-identical.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256)
    tnf_peer_adam_multimem_kernel(const __grid_constant__ PeerAdamArgs A, const float* __restrict__ grads_mc,
                                  float* __restrict__ params_mc) {
  const long long n4 = (A.shard_end - A.shard_begin) >> 2;
  const int me = A.a.rank;
  float4* __restrict__ M = reinterpret_cast<float4*>(A.m);
  float4* __restrict__ V = reinterpret_cast<float4*>(A.v);
  constexpr int U = 4;  // switch reductions in flight per thread: the round trip through the NVSwitch is long
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
    float4 g[U];
    int seg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      seg[u] = -1;
      if (i >= n4) continue;
      const long long e = A.shard_begin + 4 * i;
      int s = 0;
      while (s + 1 < A.nseg && e >= A.seg_end[s]) ++s;
      if (!A.seg_active[s]) continue;
      seg[u] = s;
      g[u] = multimem_ld_reduce_add(grads_mc + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (seg[u] < 0) continue;
      const long long i = i0 + u * stride;
      const long long e = A.shard_begin + 4 * i;
      float4 gg = g[u];
      gg.x *= A.inv_world; gg.y *= A.inv_world; gg.z *= A.inv_world; gg.w *= A.inv_world;
      float4 p = *reinterpret_cast<const float4*>(A.a.params[me] + e);
      float4 m = M[i], v = V[i];
      // never-touched entries: the update is exactly zero and every rank already holds the same value
      if (gg.x == 0.f && gg.y == 0.f && gg.z == 0.f && gg.w == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f &&
          m.w == 0.f && v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f)
        continue;
      const float ss = A.seg_step_size[seg[u]], ib = A.seg_inv_sqrt_bc2[seg[u]];
      peer_adam_one(p.x, gg.x, m.x, v.x, A, ss, ib);
      peer_adam_one(p.y, gg.y, m.y, v.y, A, ss, ib);
      peer_adam_one(p.z, gg.z, m.z, v.z, A, ss, ib);
      peer_adam_one(p.w, gg.w, m.w, v.w, A, ss, ib);
      M[i] = m;
      V[i] = v;
      multimem_st(params_mc + e, p);
    }
  }
}

// all-gather, pull flavour: every rank copies the other ranks' freshly updated shards out of their owners'
// parameter arenas (peer loads run at the NVLink rate; see DESIGN.md for the measured push / pull comparison)
__global__ void __launch_bounds__(256) tnf_peer_gather_kernel(const __grid_constant__ TnfPeerArena a,
                                                              const __grid_constant__ PeerAdamArgs A) {
  const int W = a.world_size, me = a.rank;
  const long long shard4 = (a.numel / W) >> 2;
  // 4 KB chunks dealt round-robin over the owners, rotated by the rank: at any moment the CTAs of one GPU pull from
  // all owners and the GPUs pull from different owners (walking the owners one after the other makes every rank
  // read the same GPU at the same time and divides that GPU's egress by N-1)
  const long long chunks_per_owner = (shard4 + 255) / 256, total_chunks = chunks_per_owner * (W - 1);
  for (long long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int slot = (int)((c + me) % (W - 1));
    const int owner = slot >= me ? slot + 1 : slot;  // skip the own shard
    const long long off4 = (c / (W - 1)) * 256 + threadIdx.x;
    if (off4 >= shard4) continue;
    const long long e = ((long long)owner * shard4 + off4) << 2;
    int s = 0;
    while (s + 1 < A.nseg && e >= A.seg_end[s]) ++s;
    if (!A.seg_active[s]) continue;
    *reinterpret_cast<float4*>(a.params[me] + e) = __ldcv(reinterpret_cast<const float4*>(a.params[owner] + e));
  }
}

// Flag barrier over peer-mapped words: rank r writes `epoch` into word [slot][r] of every rank's flag block
// (release, system scope) and waits until every word of its own block reached `epoch` (acquire).
// The wait is bounded (`limit` SM cycles; TNF_PEER_TIMEOUT_S seconds, default 60): a peer that never arrives is
// FATAL - word [TNF_PEER_FLAG_TIMEOUT] is raised and the kernel traps, so the next CUDA call of this rank fails
// instead of the step carrying on with stale or zeroed peer gradients (ranks would diverge silently).
__global__ void tnf_peer_barrier_kernel(const __grid_constant__ TnfPeerArena a, const int slot, const unsigned epoch,
                                        const long long limit) {
  const int t = threadIdx.x;
  if (t >= a.world_size) return;
  __threadfence_system();
  unsigned* dst = a.flags[t] + slot * TNF_MAX_PEERS + a.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  const unsigned* src = a.flags[a.rank] + slot * TNF_MAX_PEERS + t;
  const long long t0 = clock64();
  unsigned seen = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(src) : "memory");
    if ((int)(seen - epoch) >= 0) break;
    if (clock64() - t0 > limit) {
      atomicAdd(a.flags[a.rank] + TNF_PEER_FLAG_TIMEOUT, 1u);
      __threadfence_system();
      __trap();
    }
  }
  __threadfence_system();
}

}  // namespace tnf

extern "C" {

int tnf_peer_enable_access(int32_t peer_device) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (peer_device == dev) return TNF_OK;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaDeviceCanAccessPeer: %s", cudaGetErrorString(e));
  if (!can) return fail(TNF_ERR_UNSUPPORTED_CONFIG, "device %d cannot access device %d (no NVLink/PCIe P2P)", dev,
                        peer_device);
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-free error
    return TNF_OK;
  }
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
  return TNF_OK;
}

int tnf_peer_alloc(size_t bytes, void** device_ptr, void* ipc_handle) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!device_ptr || !ipc_handle || bytes == 0) return fail(TNF_ERR_INVALID_ARGUMENT, "device_ptr/ipc_handle null or bytes == 0");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);  // its own allocation: cudaIpcGetMemHandle wants the base pointer
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return fail(TNF_ERR_CUDA, "cudaMemset/cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle, &h, sizeof(h));
  *device_ptr = p;
  return TNF_OK;
}

int tnf_peer_free(void* device_ptr) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!device_ptr) return TNF_OK;
  const cudaError_t e = cudaFree(device_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TNF_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e));
  }
  return TNF_OK;
}

int tnf_peer_open_handle(const void* ipc_handle, void** device_ptr) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!ipc_handle || !device_ptr) return fail(TNF_ERR_INVALID_ARGUMENT, "ipc_handle/device_ptr is null");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == TNF_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(&h, ipc_handle, sizeof(h));
  // opened with the CONSUMER device current: the mapping (and, for a remote GPU, peer access) is created
  // for the device whose kernels will dereference the pointer
  const cudaError_t e = cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TNF_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  return TNF_OK;
}

int tnf_peer_close_handle(void* device_ptr) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!device_ptr) return TNF_OK;
  const cudaError_t e = cudaIpcCloseMemHandle(device_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TNF_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  }
  return TNF_OK;
}

static int check_arena(const TnfPeerArena* a) {
  using tnf::fail;
  if (!a) return fail(TNF_ERR_INVALID_ARGUMENT, "arena is null");
  if (a->world_size < 1 || a->world_size > TNF_MAX_PEERS || a->rank < 0 || a->rank >= a->world_size)
    return fail(TNF_ERR_INVALID_ARGUMENT, "world_size=%d rank=%d (at most %d peers)", a->world_size, a->rank,
                TNF_MAX_PEERS);
  for (int r = 0; r < a->world_size; ++r)
    if (!a->flags[r]) return fail(TNF_ERR_INVALID_ARGUMENT, "flags[%d] is null", r);
  return TNF_OK;
}

int tnf_peer_barrier(const TnfPeerArena* arena, int32_t slot, uint32_t epoch, void* stream_) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (int rc = check_arena(arena)) return rc;
  if (slot < 0 || slot >= TNF_PEER_FLAG_SLOTS) return fail(TNF_ERR_INVALID_ARGUMENT, "slot=%d", slot);
  static long long limit = 0;
  if (limit == 0) {  // seconds of SM clock at ~2 GHz; generous: a rank may be saving a checkpoint or evaluating
    const char* v = getenv("TNF_PEER_TIMEOUT_S");
    const double sec = v ? atof(v) : 60.0;
    limit = (long long)((sec > 0.01 ? sec : 0.01) * 2.0e9);
  }
  tnf::tnf_peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(*arena, slot, epoch, limit);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "peer barrier launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

static int peer_adam_impl(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                          const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                          int phase, void* stream_, const float* grads_mc = nullptr, float* params_mc = nullptr,
                          long long range_begin = 0, long long range_end = -1, int max_ctas = 0);

int tnf_peer_adam_multimem(const TnfPeerArena* arena, const float* grads_multicast, float* params_multicast,
                           float* exp_avg_shard, float* exp_avg_sq_shard, const TnfAdamSegment* segments,
                           int32_t num_segments, double beta1, double beta2, float eps, void* stream_) {
  if (!grads_multicast || !params_multicast || !tnf::aligned16(grads_multicast) || !tnf::aligned16(params_multicast)) {
    tnf::g_err[0] = 0;
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "multicast addresses null or not 16-byte aligned");
  }
  return peer_adam_impl(arena, exp_avg_shard, exp_avg_sq_shard, segments, num_segments, beta1, beta2, eps, 3, stream_,
                        grads_multicast, params_multicast);
}

int tnf_peer_adam_step(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                       const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                       void* stream_) {
  return peer_adam_impl(arena, exp_avg_shard, exp_avg_sq_shard, segments, num_segments, beta1, beta2, eps, 0, stream_);
}

int tnf_peer_adam_reduce(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                         const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                         void* stream_) {
  return peer_adam_impl(arena, exp_avg_shard, exp_avg_sq_shard, segments, num_segments, beta1, beta2, eps, 1, stream_);
}

int tnf_peer_adam_range(const TnfPeerArena* arena, int32_t flavour, int64_t range_begin, int64_t range_end,
                        int32_t max_ctas, const float* grads_multicast, float* params_multicast,
                        float* exp_avg_shard, float* exp_avg_sq_shard, const TnfAdamSegment* segments,
                        int32_t num_segments, double beta1, double beta2, float eps, void* stream_) {
  tnf::g_err[0] = 0;
  if (flavour != TNF_PEER_PUSH && flavour != TNF_PEER_MULTIMEM)
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "flavour=%d (TNF_PEER_PUSH or TNF_PEER_MULTIMEM)", flavour);
  if (flavour == TNF_PEER_MULTIMEM && (!grads_multicast || !params_multicast || !tnf::aligned16(grads_multicast) ||
                                       !tnf::aligned16(params_multicast)))
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "multicast addresses null or not 16-byte aligned");
  if (range_end < range_begin || range_begin < 0)
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "range [%lld,%lld)", (long long)range_begin, (long long)range_end);
  return peer_adam_impl(arena, exp_avg_shard, exp_avg_sq_shard, segments, num_segments, beta1, beta2, eps, flavour,
                        stream_, grads_multicast, params_multicast, range_begin, range_end, max_ctas);
}

int tnf_peer_gather_params(const TnfPeerArena* arena, const TnfAdamSegment* segments, int32_t num_segments,
                           void* stream_) {
  return peer_adam_impl(arena, nullptr, nullptr, segments, num_segments, 0.9, 0.999, 0.f, 2, stream_);
}

// phase 0: reduce + Adam + push to all ranks; 1: reduce + Adam, own copy only; 2: pull the other ranks' shards;
// 3: reduce + Adam + broadcast through the NVSwitch multicast objects
static int peer_adam_impl(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                          const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                          int phase, void* stream_, const float* grads_mc, float* params_mc, long long range_begin,
                          long long range_end, int max_ctas) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (int rc = check_arena(arena)) return rc;
  const int W = arena->world_size;
  if (arena->numel < 0 || arena->numel % (4LL * W) != 0)
    return fail(TNF_ERR_INVALID_ARGUMENT, "arena numel=%lld must be a multiple of 4*world_size", (long long)arena->numel);
  for (int r = 0; r < W; ++r)
    if (!arena->grads[r] || !arena->params[r] || !tnf::aligned16(arena->grads[r]) || !tnf::aligned16(arena->params[r]))
      return fail(TNF_ERR_INVALID_ARGUMENT, "grads[%d]/params[%d] null or not 16-byte aligned", r, r);
  if (phase != 2 &&
      (!exp_avg_shard || !exp_avg_sq_shard || !tnf::aligned16(exp_avg_shard) || !tnf::aligned16(exp_avg_sq_shard)))
    return fail(TNF_ERR_INVALID_ARGUMENT, "exp_avg/exp_avg_sq shard null or not 16-byte aligned");
  if (!segments || num_segments < 1 || num_segments > TNF_MAX_ADAM_SEGMENTS)
    return fail(TNF_ERR_INVALID_ARGUMENT, "num_segments=%d not in [1,%d]", num_segments, TNF_MAX_ADAM_SEGMENTS);
  tnf::PeerAdamArgs A;
  A.a = *arena;
  if (range_end < 0) range_end = arena->numel;
  if (range_begin < 0 || range_end > arena->numel || range_begin > range_end || (range_begin & 3) ||
      (range_end - range_begin) % (4LL * W) != 0)
    return fail(TNF_ERR_INVALID_ARGUMENT, "range [%lld,%lld): must lie in the arena, start on a multiple of 4 and "
                "have a length that is a multiple of 4*world_size", range_begin, range_end);
  if (phase == 2 && (range_begin != 0 || range_end != arena->numel))
    return fail(TNF_ERR_INVALID_ARGUMENT, "the pull all-gather works on the whole arena");
  const long long shard = (range_end - range_begin) / W;
  A.shard_begin = range_begin + shard * arena->rank;
  A.shard_end = A.shard_begin + shard;
  A.m = exp_avg_shard;
  A.v = exp_avg_sq_shard;
  long long prev = 0;
  for (int s = 0; s < num_segments; ++s) {
    const TnfAdamSegment& g = segments[s];
    if (g.begin != prev || g.end < g.begin || (g.end & 3))
      return fail(TNF_ERR_INVALID_ARGUMENT, "segment %d: [%lld,%lld) must continue the previous one and end on a "
                  "multiple of 4", s, (long long)g.begin, (long long)g.end);
    if (g.active && g.step < 1) return fail(TNF_ERR_INVALID_ARGUMENT, "segment %d: step=%lld must be >= 1", s,
                                           (long long)g.step);
    prev = g.end;
    A.seg_end[s] = g.end;
    A.seg_active[s] = g.active != 0;
    const double st = g.active ? (double)g.step : 1.0;
    A.seg_step_size[s] = (float)((double)g.lr / (1.0 - pow(beta1, st)));
    A.seg_inv_sqrt_bc2[s] = 1.0f / (float)sqrt(1.0 - pow(beta2, st));
  }
  if (prev != arena->numel) return fail(TNF_ERR_INVALID_ARGUMENT, "segments cover %lld of %lld elements", prev,
                                        (long long)arena->numel);
  A.nseg = num_segments;
  A.omb1 = (float)(1.0 - beta1);
  A.omb2 = (float)(1.0 - beta2);
  A.beta2 = (float)beta2;
  A.eps = eps;
  A.inv_world = 1.0f / (float)W;
  A.push = phase == 0;
  const long long n4 = phase == 2 ? (((shard >> 2) + 255) / 256) * 256 * (W - 1) : (shard >> 2);
  if (n4 == 0) return TNF_OK;
  long long blocks = (n4 + 255) / 256;
  const long long cap = max_ctas > 0 ? (long long)max_ctas : (long long)tnf::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (phase == 2)
    tnf::tnf_peer_gather_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(*arena, A);
  else if (phase == 3)
    tnf::tnf_peer_adam_multimem_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(A, grads_mc,
                                                                                                       params_mc);
  else
    tnf::tnf_peer_adam_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(A);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "peer adam launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

}  // extern "C"
