// Device-side building blocks shared by the forward and backward kernels of
// libtnf_b200.so.  Everything here is warp-level: one warp owns one ray.
//
// Math follows the nerfstudio-1.1.5 torch implementation reached from
// thermo_nerf/thermal_nerf/thermal_nerf_model.py:210-275;
// the citations on each helper name the reference line that reaches it.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/tnf_b200.h"

namespace tnf {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kBuf = 264;        // floats per per-warp scratch line (>= TNF_MAX_SAMPLES + 1)
constexpr int kMaxFieldS = 64;   // max samples of the final (field) level
constexpr uint32_t kPrimeY = 2654435761u;
constexpr uint32_t kPrimeZ = 805459861u;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------- small math
__device__ __forceinline__ float nan_to_num(float x) {
  // torch.nan_to_num defaults: nan -> 0, +-inf -> +-FLT_MAX
  if (isnan(x)) return 0.f;
  if (isinf(x)) return copysignf(FLT_MAX, x);
  return x;
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }
// Tensor-core mode only: sigmoid(x) = 0.5 tanh(x/2) + 0.5 on the MUFU tanh unit (one SFU op instead of
// exp + full-precision divide; |error| <= 2.5e-4, below the fp16 rounding of the activations it feeds).
__device__ __forceinline__ float sigmoid_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// UniformLinDispPiecewiseSampler spacing functions (thermal_nerf_model.py:172-179 builds the
// ProposalNetworkSampler whose default initial sampler this is).
__device__ __forceinline__ float spacing_fn(float x) { return x < 1.f ? x * 0.5f : 1.f - 1.f / (2.f * x); }
__device__ __forceinline__ float spacing_inv(float x) { return x < 0.5f ? 2.f * x : 1.f / (2.f - 2.f * x); }
__device__ __forceinline__ float to_euclid(float s, float s_near, float s_far) {
  return spacing_inv(s * s_far + (1.f - s) * s_near);
}

// torch.linspace element (same two-sided evaluation as ATen's kernel).
__device__ __forceinline__ float linspace_at(int i, int steps, float start, float end) {
  const float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

// Spacing bin i (0..S) of the initial piecewise sampler; `jit` is the single-jitter draw.
__device__ __forceinline__ float initial_sbin(int i, int S, bool stratified, float jit) {
  const float b = linspace_at(i, S + 1, 0.f, 1.f);
  if (!stratified) return b;
  const float lo = (i == 0) ? b : (b + linspace_at(i - 1, S + 1, 0.f, 1.f)) * 0.5f;
  const float hi = (i == S) ? b : (linspace_at(i + 1, S + 1, 0.f, 1.f) + b) * 0.5f;
  return lo + (hi - lo) * jit;
}

// ---------------------------------------------------------------- warp collectives
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ---------------------------------------------------------------- positions
// Shared head of NerfactoField.get_density / HashMLPDensityField.get_density:
// L-inf contraction -> (x+2)/4 -> selector -> zero out-of-range points.
__device__ __forceinline__ float normalise_position(const TnfModel& m, float x, float y, float z, float& px,
                                                    float& py, float& pz) {
  if (m.use_contraction) {
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    if (!(mag < 1.f)) {
      const float s = 2.f - 1.f / mag;
      x = s * (x / mag);
      y = s * (y / mag);
      z = s * (z / mag);
    }
    px = (x + 2.f) * 0.25f;
    py = (y + 2.f) * 0.25f;
    pz = (z + 2.f) * 0.25f;
  } else {
    px = (x - m.aabb[0]) / (m.aabb[3] - m.aabb[0]);
    py = (y - m.aabb[1]) / (m.aabb[4] - m.aabb[1]);
    pz = (z - m.aabb[2]) / (m.aabb[5] - m.aabb[2]);
  }
  const bool sel = (px > 0.f) && (px < 1.f) && (py > 0.f) && (py < 1.f) && (pz > 0.f) && (pz < 1.f);
  const float s = sel ? 1.f : 0.f;
  px *= s;
  py *= s;
  pz *= s;
  return s;
}

// ---------------------------------------------------------------- hash grid
// One level of HashEncoding.pytorch_fwd.  Coordinates are >= 0 and the table size is a
// power of two, so the reference's int64 product/xor/mod equals 32-bit mul.lo/xor/and
// bit for bit.  `level_tab` already points at the level's first entry.
struct HashCorners {
  uint32_t idx[8];
  float ox, oy, oz;
  uint32_t k1, k2;  // exact identity of the cell (floor coordinates + which axes have ceil != floor)
};

__device__ __forceinline__ void hash_corners(float px, float py, float pz, float scale, uint32_t mask,
                                             HashCorners& hc) {
  const float sx = px * scale, sy = py * scale, sz = pz * scale;
  const float fxf = floorf(sx), fyf = floorf(sy), fzf = floorf(sz);
  const float cxf = ceilf(sx), cyf = ceilf(sy), czf = ceilf(sz);
  hc.ox = sx - fxf;
  hc.oy = sy - fyf;
  hc.oz = sz - fzf;
  const uint32_t fx = (uint32_t)(int)fxf, cx = (uint32_t)(int)cxf;
  const uint32_t fy = (uint32_t)(int)fyf * kPrimeY, cy = (uint32_t)(int)cyf * kPrimeY;
  const uint32_t fz = (uint32_t)(int)fzf * kPrimeZ, cz = (uint32_t)(int)czf * kPrimeZ;
  hc.k1 = fx | ((uint32_t)(int)fyf << 11) | ((cx != fx ? 1u : 0u) << 22) | ((cyf != fyf ? 1u : 0u) << 23) |
          ((czf != fzf ? 1u : 0u) << 24);
  hc.k2 = (uint32_t)(int)fzf;
  hc.idx[0] = (cx ^ cy ^ cz) & mask;
  hc.idx[1] = (cx ^ fy ^ cz) & mask;
  hc.idx[2] = (fx ^ fy ^ cz) & mask;
  hc.idx[3] = (fx ^ cy ^ cz) & mask;
  hc.idx[4] = (cx ^ cy ^ fz) & mask;
  hc.idx[5] = (cx ^ fy ^ fz) & mask;
  hc.idx[6] = (fx ^ fy ^ fz) & mask;
  hc.idx[7] = (fx ^ cy ^ fz) & mask;
}

__device__ __forceinline__ float2 hash_blend(const float2 (&f)[8], float ox, float oy, float oz) {
  const float ix = 1.f - ox, iy = 1.f - oy, iz = 1.f - oz;
  float2 f03, f12, f56, f47, a, b, o;
  f03.x = f[0].x * ox + f[3].x * ix;  f03.y = f[0].y * ox + f[3].y * ix;
  f12.x = f[1].x * ox + f[2].x * ix;  f12.y = f[1].y * ox + f[2].y * ix;
  f56.x = f[5].x * ox + f[6].x * ix;  f56.y = f[5].y * ox + f[6].y * ix;
  f47.x = f[4].x * ox + f[7].x * ix;  f47.y = f[4].y * ox + f[7].y * ix;
  a.x = f03.x * oy + f12.x * iy;      a.y = f03.y * oy + f12.y * iy;
  b.x = f47.x * oy + f56.x * iy;      b.y = f47.y * oy + f56.y * iy;
  o.x = a.x * oz + b.x * iz;          o.y = a.y * oz + b.y * iz;
  return o;
}

__device__ __forceinline__ float2 hash_level(const float2* __restrict__ level_tab, float px, float py, float pz,
                                             float scale, uint32_t mask) {
  HashCorners hc;
  hash_corners(px, py, pz, scale, mask, hc);
  float2 f[8];
  // (Tried in round 2: fetching the two x-neighbours of an even floor(x) - table entries e and e ^ 1 - with one aligned
  // 16-byte load and issuing the second 8-byte gather only from odd lanes.  Bit-identical, but 9 % SLOWER on the 800x800
  // frame and on the training forward: the selects and the divergent second gather cost more than the wavefronts saved.)
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = __ldg(level_tab + hc.idx[c]);
  return hash_blend(f, hc.ox, hc.oy, hc.oz);
}

// ---------------------------------------------------------------- SH (degree 4)
// components_from_spherical_harmonics evaluated on (d+1)/2 (thermal_field.py:117-119).
__device__ __forceinline__ void sh4(float x, float y, float z, float (&c)[16]) {
  const float xx = x * x, yy = y * y, zz = z * z;
  c[0] = 0.28209479177387814f;
  c[1] = 0.4886025119029199f * y;
  c[2] = 0.4886025119029199f * z;
  c[3] = 0.4886025119029199f * x;
  c[4] = 1.0925484305920792f * x * y;
  c[5] = 1.0925484305920792f * y * z;
  c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
  c[7] = 1.0925484305920792f * x * z;
  c[8] = 0.5462742152960396f * (xx - yy);
  c[9] = 0.5900435899266435f * y * (3.f * xx - yy);
  c[10] = 2.890611442640554f * x * y * z;
  c[11] = 0.4570457994644658f * y * (5.f * zz - 1.f);
  c[12] = 0.3731763325901154f * z * (5.f * zz - 3.f);
  c[13] = 0.4570457994644658f * x * (5.f * zz - 1.f);
  c[14] = 1.445305721320277f * z * (xx - yy);
  c[15] = 0.5900435899266435f * x * (xx - 3.f * yy);
}

// ---------------------------------------------------------------- tensor-core helpers
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// four 8x8 b16 matrices from shared memory in mma A-fragment order (row-major source)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// ---------------------------------------------------------------- shared-memory images
// Proposal density MLP (grid -> 16 -> 1), fp32, k-major so one float4 feeds 4 FMAs.
struct PropW {
  float w0t[16 * 16];  // [k][j], k < 2*levels (rest zero)
  float b0[16];
  float w1[16];
  float b1;
  float pad[3];
};

// Per-warp scratch: one ray in flight per warp.
struct WarpScratch {
  float w[kBuf];     // weights of the current level, later the next level's spacing bins
  float cdf[kBuf];   // cdf of the current level
  float bins[kBuf];  // spacing bins of the current level
  float sigma[kMaxFieldS];
  float r[kMaxFieldS];
  float g[kMaxFieldS];
  float b[kMaxFieldS];
  float th[kMaxFieldS];
  float rayb[64];  // per-ray first-layer bias of the colour head (bias + SH + appearance part)
};

}  // namespace tnf
