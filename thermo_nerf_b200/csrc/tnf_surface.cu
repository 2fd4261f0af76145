// The Field / Renderer plugin surface of the path as stand-alone sm_100a kernels:
//
//   tnf_field_density   NerfactoField.get_density reached from ThermalNerfactoTField.forward
//                       (thermo_nerf/thermal_nerf/thermal_field.py:183-201) and HashMLPDensityField.density_fn
//                       (proposal networks, thermal_nerf_model.py:127-148) on N arbitrary points
//   tnf_field_heads     ThermalNerfactoTField.get_outputs (thermal_field.py:108-181): SH4((d+1)/2), appearance
//                       embedding, colour head 63-64-64-3, thermal head 15-64-64-1 from the 15 geo features
//   tnf_composite       ThermalRenderer.forward (thermal_renderer.py:113-149; background always the last sample,
//                       :49,68-70) and RGBTRenderer.forward of the concat baseline (rgb_concat/rgbt_renderer.py:134-140,
//                       no background term) on [R,S,C] per-sample values
//
// These are the callers' per-module entry points (viewer density queries, user code that composes the modules
// by hand); the training / render hot path is the fused kernel of tnf_forward.cu.  Thread per sample, fp32
// arithmetic in the reference's operation order, weights read through the read-only path (every lane of a warp
// reads the same address: one transaction per load).
#include "tnf_field.cuh"
#include "tnf_host.h"

namespace tnf {

__device__ __forceinline__ float lin_row(const float* __restrict__ w, const float* __restrict__ b, const int n,
                                         const float* x, const int K, const int ld) {
  float acc = __ldg(b + n);
  for (int k = 0; k < K; ++k) acc = fmaf(x[k], __ldg(w + n * ld + k), acc);
  return acc;
}

__global__ void __launch_bounds__(128)
    tnf_field_density_kernel(const __grid_constant__ TnfModel m, const int which, const float* __restrict__ pos,
                             const long long n, float* __restrict__ density, float* __restrict__ geo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px, py, pz;
  const float sel = normalise_position(m, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], px, py, pz);
  if (which >= 0) {  // HashMLPDensityField: grid -> 16 -> 1
    const TnfDensityNet& net = m.prop[which];
    const int L = net.grid.num_levels;
    const uint32_t mask = (1u << net.grid.log2_size) - 1u;
    const float2* tab = reinterpret_cast<const float2*>(net.grid.table);
    float feat[2 * TNF_MAX_PROP_LEVELS];
    for (int l = 0; l < L; ++l) {
      const float2 f = hash_level(tab + ((size_t)l << net.grid.log2_size), px, py, pz, net.grid.scalings[l], mask);
      feat[2 * l] = f.x;
      feat[2 * l + 1] = f.y;
    }
    float o = __ldg(net.l1.bias);
    for (int j = 0; j < 16; ++j) {
      const float h = lin_row(net.l0.weight, net.l0.bias, j, feat, 2 * L, 2 * L);
      o = fmaf(fmaxf(h, 0.f), __ldg(net.l1.weight + j), o);
    }
    density[i] = expf(o) * sel;
    return;
  }
  const TnfField& f = m.field;
  const uint32_t mask = (1u << f.grid.log2_size) - 1u;
  const float2* tab = reinterpret_cast<const float2*>(f.grid.table);
  float feat[32];
#pragma unroll 4
  for (int l = 0; l < TNF_MAX_LEVELS; ++l) {
    const float2 v = hash_level(tab + ((size_t)l << f.grid.log2_size), px, py, pz, f.grid.scalings[l], mask);
    feat[2 * l] = v.x;
    feat[2 * l + 1] = v.y;
  }
  float h[64];
  for (int j = 0; j < 64; ++j) h[j] = fmaxf(lin_row(f.base0.weight, f.base0.bias, j, feat, 32, 32), 0.f);
  for (int j = 0; j < 16; ++j) {
    const float o = lin_row(f.base1.weight, f.base1.bias, j, h, 64, 64);
    if (j == 0) density[i] = expf(o) * sel;
    else if (geo) geo[15 * i + j - 1] = o;
  }
}

__global__ void __launch_bounds__(128)
    tnf_field_heads_kernel(const __grid_constant__ TnfModel m, const float* __restrict__ dirs,
                           const long long* __restrict__ cam, const float* __restrict__ geo, const long long n,
                           float* __restrict__ rgb, float* __restrict__ thermal) {
  __shared__ float app_const[32];
  if (threadIdx.x < 32) {
    float e = 0.f;
    if (m.appearance_mode == TNF_APPEARANCE_MEAN) {
      for (int k = 0; k < m.field.num_images; ++k) e += m.field.appearance[k * 32 + threadIdx.x];
      e /= (float)m.field.num_images;
    }
    app_const[threadIdx.x] = e;
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TnfField& f = m.field;
  float x[63];
  {
    float sh[16];
    sh4((dirs[3 * i] + 1.f) * 0.5f, (dirs[3 * i + 1] + 1.f) * 0.5f, (dirs[3 * i + 2] + 1.f) * 0.5f, sh);
    for (int k = 0; k < 16; ++k) x[k] = sh[k];
  }
  for (int k = 0; k < 15; ++k) x[16 + k] = geo[15 * i + k];
  if (m.appearance_mode == TNF_APPEARANCE_LOOKUP) {
    const long long c = cam[i];
    for (int k = 0; k < 32; ++k) x[31 + k] = __ldg(f.appearance + c * 32 + k);
  } else {
    for (int k = 0; k < 32; ++k) x[31 + k] = app_const[k];
  }
  float a[64], b[64];
  if (rgb) {
    for (int j = 0; j < 64; ++j) a[j] = fmaxf(lin_row(f.rgb0.weight, f.rgb0.bias, j, x, 63, 63), 0.f);
    for (int j = 0; j < 64; ++j) b[j] = fmaxf(lin_row(f.rgb1.weight, f.rgb1.bias, j, a, 64, 64), 0.f);
    for (int j = 0; j < 3; ++j) rgb[3 * i + j] = sigmoidf(lin_row(f.rgb2.weight, f.rgb2.bias, j, b, 64, 64));
  }
  if (thermal) {
    for (int j = 0; j < 64; ++j) a[j] = fmaxf(lin_row(f.th0.weight, f.th0.bias, j, x + 16, 15, 15), 0.f);
    for (int j = 0; j < 64; ++j) b[j] = sigmoidf(lin_row(f.th1.weight, f.th1.bias, j, a, 64, 64));
    thermal[i] = lin_row(f.th2.weight, f.th2.bias, 0, b, 64, 64);
  }
}

// out[r][c] = sum_s w[r][s] v[r][s][c]  (+ v[r][S-1][c] (1 - sum_s w[r][s]) for the last-sample background);
// eval: nan_to_num on the samples, clamp to [0,1] on the result.  One warp per ray.
__global__ void __launch_bounds__(256)
    tnf_composite_kernel(const float* __restrict__ values, const float* __restrict__ weights, const long long R,
                         const int S, const int C, const int last_sample_bg, const int eval_mode,
                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= R) return;
  const float* w = weights + ray * S;
  const float* v = values + ray * (long long)S * C;
  float sw = 0.f;
  for (int s = lane; s < S; s += 32) sw += w[s];
  sw = warp_sum(sw);
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int s = lane; s < S; s += 32) {
      float x = v[(long long)s * C + c];
      if (eval_mode) x = nan_to_num(x);
      acc = fmaf(w[s], x, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (last_sample_bg) {
        float bg = v[(long long)(S - 1) * C + c];
        if (eval_mode) bg = nan_to_num(bg);
        acc += bg * (1.f - sw);
      }
      if (eval_mode) acc = fminf(fmaxf(acc, 0.f), 1.f);
      out[ray * C + c] = acc;
    }
  }
}

}  // namespace tnf

namespace {
using tnf::fail;

// the surface entry points use one part of TnfModel each: only that part has to be filled in
int check_density_net(const TnfDensityNet& n, int k) {
  if (!n.grid.table || !n.l0.weight || !n.l0.bias || !n.l1.weight || !n.l1.bias)
    return fail(TNF_ERR_INVALID_ARGUMENT, "prop[%d]: null table / weight / bias", k);
  if (n.grid.num_levels < 1 || n.grid.num_levels > TNF_MAX_PROP_LEVELS || n.grid.log2_size < 1 || n.grid.log2_size > 24)
    return fail(TNF_ERR_UNSUPPORTED_CONFIG, "prop[%d]: num_levels=%d log2_size=%d", k, n.grid.num_levels,
                n.grid.log2_size);
  return TNF_OK;
}
int check_field(const TnfModel& m, bool need_grid, bool need_heads) {
  const TnfField& f = m.field;
  if (need_grid) {
    if (!f.grid.table || !f.base0.weight || !f.base0.bias || !f.base1.weight || !f.base1.bias)
      return fail(TNF_ERR_INVALID_ARGUMENT, "field: null table / mlp_base weight");
    if (f.grid.num_levels != TNF_MAX_LEVELS || f.grid.log2_size < 1 || f.grid.log2_size > 24)
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "field: num_levels=%d log2_size=%d", f.grid.num_levels, f.grid.log2_size);
  }
  if (need_heads) {
    const TnfLinear* ls[] = {&f.rgb0, &f.rgb1, &f.rgb2, &f.th0, &f.th1, &f.th2};
    for (const TnfLinear* l : ls)
      if (!l->weight || !l->bias) return fail(TNF_ERR_INVALID_ARGUMENT, "field: a head weight / bias is null");
    if (m.appearance_mode < 0 || m.appearance_mode > 2) return fail(TNF_ERR_INVALID_ARGUMENT, "appearance_mode=%d", m.appearance_mode);
    if (m.appearance_mode != TNF_APPEARANCE_ZEROS && (!f.appearance || f.num_images < 1))
      return fail(TNF_ERR_INVALID_ARGUMENT, "field.appearance is required for appearance_mode=%d", m.appearance_mode);
  }
  return TNF_OK;
}
}  // namespace

extern "C" {

int tnf_field_density(const TnfModel* model, int32_t which, const float* positions, int64_t n, float* density,
                      float* geo, void* stream_) {
  tnf::g_err[0] = 0;
  if (!model) return fail(TNF_ERR_INVALID_ARGUMENT, "model is null");
  if (which < -1 || which >= TNF_NUM_PROP) return fail(TNF_ERR_INVALID_ARGUMENT, "which=%d not in [-1,%d)", which, TNF_NUM_PROP);
  if (int e = which >= 0 ? check_density_net(model->prop[which], which) : check_field(*model, true, false)) return e;
  if (n < 0) return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "n=%lld", (long long)n);
  if (n == 0) return TNF_OK;
  if (!positions || !density) return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "positions/density is null");
  const int tb = 128;
  tnf::tnf_field_density_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream_)>>>(
      *model, which, positions, n, density, which < 0 ? geo : nullptr);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tnf::fail(TNF_ERR_CUDA, "field_density launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

int tnf_field_heads(const TnfModel* model, const float* directions, const int64_t* camera_indices, const float* geo,
                    int64_t n, float* rgb, float* thermal, void* stream_) {
  tnf::g_err[0] = 0;
  if (!model) return fail(TNF_ERR_INVALID_ARGUMENT, "model is null");
  if (int e = check_field(*model, false, true)) return e;
  if (n < 0) return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "n=%lld", (long long)n);
  if (n == 0) return TNF_OK;
  if (!directions || !geo) return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "directions/geo is null");
  if (model->appearance_mode == TNF_APPEARANCE_LOOKUP && !camera_indices)
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "camera_indices required for TNF_APPEARANCE_LOOKUP");
  const int tb = 128;
  tnf::tnf_field_heads_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream_)>>>(
      *model, directions, reinterpret_cast<const long long*>(camera_indices), geo, n, rgb, thermal);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tnf::fail(TNF_ERR_CUDA, "field_heads launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

int tnf_composite(const float* values, const float* weights, int64_t num_rays, int32_t num_samples,
                  int32_t channels, int32_t last_sample_background, int32_t eval_mode, float* out, void* stream_) {
  tnf::g_err[0] = 0;
  if (num_rays < 0 || num_samples < 1 || channels < 1)
    return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "num_rays=%lld num_samples=%d channels=%d", (long long)num_rays,
                     num_samples, channels);
  if (num_rays == 0) return TNF_OK;
  if (!values || !weights || !out) return tnf::fail(TNF_ERR_INVALID_ARGUMENT, "values/weights/out is null");
  const int tb = 256;
  const long long threads = num_rays * 32;
  tnf::tnf_composite_kernel<<<(unsigned)((threads + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream_)>>>(
      values, weights, num_rays, num_samples, channels, last_sample_background, eval_mode, out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tnf::fail(TNF_ERR_CUDA, "composite launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

}  // extern "C"
