// Multi-tensor Adam for the optimisers of thermo_nerf/thermal_nerf/config_thermal_nerf.py:32-45
// (AdamOptimizerConfig(lr=1e-2, eps=1e-15) for "proposal_networks" and "fields").
// Update rule = torch.optim.Adam, amsgrad=False, weight_decay=0 (torch/optim/adam.py,
// _single_tensor_adam):  m.lerp_(g, 1-b1);  v = b2*v + (1-b2)*g*g;
//                        p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// One launch covers every tensor; HBM-bound: 16 B read + 12 B (+4 B zeroed grad) written
// per parameter.  GradScaler semantics: grad *= inv_grad_scale, skip everything if *found_inf.
#include "tnf_device.cuh"
#include "tnf_host.h"

namespace tnf {

constexpr int kAdamChunk = 4096;  // elements per CTA iteration (256 threads x 4 x float4)

struct AdamArgs {
  TnfAdamTensor t[TNF_ADAM_MAX_TENSORS];
  int chunk_start[TNF_ADAM_MAX_TENSORS + 1];  // prefix sum of ceil(numel / kAdamChunk)
  int n;
  float beta2, eps;
  float omb1, omb2;  // 1-beta1, 1-beta2 rounded from double, as torch passes them to lerp_/addcmul_
  float inv_sqrt_bc2;
  float inv_grad_scale;
  const float* grad_scale;
  const float* found_inf;
  int zero_grads;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamArgs& a, float lr,
                                         float inv_scale) {
  const float gg = g * inv_scale;
  m = m + a.omb1 * (gg - m);
  v = a.beta2 * v + a.omb2 * (gg * gg);
  const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
  p = p - lr * (m / denom);  // lr already holds step_size = lr / bias_correction1
}

__global__ void __launch_bounds__(256) tnf_adam_kernel(const __grid_constant__ AdamArgs a) {
  // GradScaler: a step with non-finite gradients is skipped, but the gradients are still consumed
  const bool skip = a.found_inf && *a.found_inf != 0.f;
  if (skip && !a.zero_grads) return;
  const float inv_scale = a.grad_scale ? a.inv_grad_scale / *a.grad_scale : a.inv_grad_scale;
  const int total = a.chunk_start[a.n];
  int ti = 0;
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    while (chunk >= a.chunk_start[ti + 1]) ++ti;  // chunks visited in increasing order per CTA
    const TnfAdamTensor& t = a.t[ti];
    const long long base = (long long)(chunk - a.chunk_start[ti]) * kAdamChunk;
    const long long n = t.numel - base < kAdamChunk ? t.numel - base : kAdamChunk;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                       reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15u) == 0;
    if (vec && n == kAdamChunk) {
      float4* P = reinterpret_cast<float4*>(t.param + base);
      float4* G = reinterpret_cast<float4*>(t.grad + base);
      float4* M = reinterpret_cast<float4*>(t.exp_avg + base);
      float4* V = reinterpret_cast<float4*>(t.exp_avg_sq + base);
      float4 p[4], g[4], m[4], v[4];
      bool untouched[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = u * 256 + threadIdx.x;
        if (skip) { G[i] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        p[u] = P[i]; g[u] = __ldcs(G + i); m[u] = M[i]; v[u] = V[i];
        untouched[u] = g[u].x == 0.f && g[u].y == 0.f && g[u].z == 0.f && g[u].w == 0.f &&
                       m[u].x == 0.f && m[u].y == 0.f && m[u].z == 0.f && m[u].w == 0.f &&
                       v[u].x == 0.f && v[u].y == 0.f && v[u].z == 0.f && v[u].w == 0.f;
      }
      if (skip) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = u * 256 + threadIdx.x;
        adam_one(p[u].x, g[u].x, m[u].x, v[u].x, a, t.lr, inv_scale);
        adam_one(p[u].y, g[u].y, m[u].y, v[u].y, a, t.lr, inv_scale);
        adam_one(p[u].z, g[u].z, m[u].z, v[u].z, a, t.lr, inv_scale);
        adam_one(p[u].w, g[u].w, m[u].w, v[u].w, a, t.lr, inv_scale);
        // entries no ray has ever touched (most of the coarse hash levels: (res+1)^3 cells in a 2^19 table) have
        // g = m = v = 0 and an update of exactly zero: skip their 12 B of stores (bit-identical result)
        const bool idle = untouched[u];
        if (!idle) { P[i] = p[u]; M[i] = m[u]; V[i] = v[u]; }
        if (a.zero_grads && !idle) G[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (long long i = threadIdx.x; i < n; i += 256) {
        if (skip) { t.grad[base + i] = 0.f; continue; }
        float p = t.param[base + i], g = t.grad[base + i], m = t.exp_avg[base + i], v = t.exp_avg_sq[base + i];
        adam_one(p, g, m, v, a, t.lr, inv_scale);
        t.param[base + i] = p; t.exp_avg[base + i] = m; t.exp_avg_sq[base + i] = v;
        if (a.zero_grads) t.grad[base + i] = 0.f;
      }
    }
  }
}

}  // namespace tnf

extern "C" int tnf_adam_step(const TnfAdamTensor* tensors, int32_t num_tensors, double beta1, double beta2, float eps,
                             int64_t step, float inv_grad_scale, const float* grad_scale, const float* found_inf,
                             int32_t zero_grads, void* stream_) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (num_tensors < 0 || num_tensors > TNF_ADAM_MAX_TENSORS)
    return fail(TNF_ERR_INVALID_ARGUMENT, "num_tensors=%d not in [0,%d]", num_tensors, TNF_ADAM_MAX_TENSORS);
  if (num_tensors == 0) return TNF_OK;
  if (!tensors) return fail(TNF_ERR_INVALID_ARGUMENT, "tensors is null");
  if (step < 1) return fail(TNF_ERR_INVALID_ARGUMENT, "step=%lld must be >= 1", (long long)step);
  tnf::AdamArgs a;
  a.n = num_tensors;
  long long chunks = 0;
  for (int i = 0; i < num_tensors; ++i) {
    const TnfAdamTensor& t = tensors[i];
    if (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq || t.numel < 0)
      return fail(TNF_ERR_INVALID_ARGUMENT, "tensor %d: null pointer or negative numel", i);
    a.t[i] = t;
    a.t[i].lr = (float)((double)t.lr / (1.0 - pow(beta1, (double)step)));
    a.chunk_start[i] = (int)chunks;
    chunks += (t.numel + tnf::kAdamChunk - 1) / tnf::kAdamChunk;
    if (chunks > 0x7fffffffLL) return fail(TNF_ERR_UNSUPPORTED_CONFIG, "too many elements for one launch");
  }
  a.chunk_start[num_tensors] = (int)chunks;
  a.omb1 = (float)(1.0 - beta1);
  a.omb2 = (float)(1.0 - beta2);
  a.beta2 = (float)beta2;
  a.eps = eps;
  const double bc2 = 1.0 - pow(beta2, (double)step);
  a.inv_sqrt_bc2 = 1.0f / (float)sqrt(bc2);
  a.inv_grad_scale = inv_grad_scale;
  a.grad_scale = grad_scale;
  a.found_inf = found_inf;
  a.zero_grads = zero_grads;
  if (chunks == 0) return TNF_OK;
  const long long cap = (long long)tnf::num_sms() * 8;
  tnf::tnf_adam_kernel<<<(unsigned)(chunks < cap ? chunks : cap), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "adam kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}
