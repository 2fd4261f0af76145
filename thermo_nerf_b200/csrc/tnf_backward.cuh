// Device helpers shared by the backward translation units (tnf_backward.cu, tnf_backward_tc.cu):
// run-merged hash-grid scatter, position gradients (camera optimiser), compositing backward, per-ray
// first-layer bias of the colour head, bf16 mma.sync wrappers.
#pragma once

#include <cuda_bf16.h>

#include "tnf_field.cuh"

namespace tnf {

// ------------------------------------------------------------------------------------
// hash-grid scatter: transpose of hash_level (same corner order / weights as hash_blend)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_level(float2* __restrict__ gtab, float px, float py, float pz, float scale,
                                              uint32_t mask, float gx, float gy) {
  if (gx == 0.f && gy == 0.f) return;
  HashCorners hc;
  hash_corners(px, py, pz, scale, mask, hc);
  const float ox = hc.ox, oy = hc.oy, oz = hc.oz, ix = 1.f - ox, iy = 1.f - oy, iz = 1.f - oz;
  const float w[8] = {ox * oy * oz, ox * iy * oz, ix * iy * oz, ix * oy * oz,
                      ox * oy * iz, ox * iy * iz, ix * iy * iz, ix * oy * iz};
#pragma unroll
  for (int c = 0; c < 8; ++c) atomicAdd(gtab + hc.idx[c], make_float2(w[c] * gx, w[c] * gy));
}

// Warp-cooperative scatter.  Lanes lane, lane+STRIDE, lane+2*STRIDE, ... hold consecutive samples of one
// ray, so at the coarser levels whole runs of them fall into the same grid cell (identical corner
// indices).  The L2 reduction units retire about one lane-address per 1.3 cycles per SM, which is what
// bounds the table scatter; each run is therefore summed with a segmented shuffle reduction first and only
// its first lane issues the 8 vector reductions.  The number of shuffle rounds adapts to the longest run
// in the warp (none at the fine levels, where every sample sits in its own cell).
// Must be called by all 32 lanes; a lane without a contribution passes gx = gy = 0.
template <int STRIDE>
__device__ __forceinline__ void scatter_level_runs(float2* __restrict__ gtab, float px, float py, float pz, float scale,
                                                   uint32_t mask, float gx, float gy, const int lane) {
  HashCorners hc;
  hash_corners(px, py, pz, scale, mask, hc);
  const uint32_t p1 = __shfl_up_sync(kFull, hc.k1, STRIDE), p2 = __shfl_up_sync(kFull, hc.k2, STRIDE);
  const bool head = lane < STRIDE || p1 != hc.k1 || p2 != hc.k2;
  const unsigned heads = __ballot_sync(kFull, head);
  const unsigned nzm = __ballot_sync(kFull, gx != 0.f || gy != 0.f);
  const unsigned cls = STRIDE == 1 ? kFull : (0x11111111u << (lane & (STRIDE - 1)));
  const unsigned later = lane + STRIDE >= 32 ? 0u : ((heads & cls) >> (lane + STRIDE));
  const int nh = later ? lane + STRIDE + __ffs(later) - 1 : 32;  // first lane of the next run of my class (or 32)
  const int len = (nh - lane + STRIDE - 1) / STRIDE;             // lanes of my class from me to the run's end
  const int maxlen = __reduce_max_sync(kFull, len);
  const float ox = hc.ox, oy = hc.oy, oz = hc.oz, ix = 1.f - ox, iy = 1.f - oy, iz = 1.f - oz;
  const float w[8] = {ox * oy * oz, ox * iy * oz, ix * iy * oz, ix * oy * oz,
                      ox * oy * iz, ox * iy * iz, ix * iy * iz, ix * oy * iz};
  float vx[8], vy[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { vx[c] = w[c] * gx; vy[c] = w[c] * gy; }
#pragma unroll
  for (int d = 1; d < 32 / STRIDE; d <<= 1) {
    if (d < maxlen) {  // warp-uniform
      const bool take = lane + d * STRIDE < nh;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float tx = __shfl_down_sync(kFull, vx[c], d * STRIDE), ty = __shfl_down_sync(kFull, vy[c], d * STRIDE);
        if (take) { vx[c] += tx; vy[c] += ty; }
      }
    }
  }
  // does any lane of my run carry a gradient?
  const unsigned runbits = (len >= 32 / STRIDE && STRIDE == 1 && lane == 0) ? kFull : 0u;
  unsigned mine = 0u;
  if (STRIDE == 1) {
    mine = runbits ? kFull : (((1u << len) - 1u) << lane);
  } else {
#pragma unroll
    for (int i = 0; i < 32 / STRIDE; ++i)
      if (i < len) mine |= 1u << (lane + i * STRIDE);
  }
  if (head && (nzm & mine)) {
#pragma unroll
    for (int c = 0; c < 8; ++c) atomicAdd(gtab + hc.idx[c], make_float2(vx[c], vy[c]));
  }
}

// ------------------------------------------------------------------------------------
// gradient w.r.t. the sample position (camera-optimiser path): d feat / d p through the trilinear weights
// (the reference differentiates offset = scaled - floor(scaled); ceil/floor carry no gradient), then back
// through (x + 2) / 4, the selector mask and the L-inf scene contraction.
// ------------------------------------------------------------------------------------
// adds scale * sum_c (g . T[idx_c]) d w_c / d offset to (dpx, dpy, dpz); corner order of hash_blend
__device__ __forceinline__ void hash_level_pos_grad(const float2* __restrict__ level_tab, const float px, const float py,
                                                    const float pz, const float scale, const uint32_t mask,
                                                    const float gx, const float gy, float& dpx, float& dpy, float& dpz) {
  HashCorners hc;
  hash_corners(px, py, pz, scale, mask, hc);
  float s[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float2 t = __ldg(level_tab + hc.idx[c]);
    s[c] = gx * t.x + gy * t.y;
  }
  const float ox = hc.ox, oy = hc.oy, oz = hc.oz, ix = 1.f - ox, iy = 1.f - oy, iz = 1.f - oz;
  dpx += scale * (oy * oz * (s[0] - s[3]) + iy * oz * (s[1] - s[2]) + oy * iz * (s[4] - s[7]) + iy * iz * (s[5] - s[6]));
  dpy += scale * (ox * oz * (s[0] - s[1]) + ix * oz * (s[3] - s[2]) + ox * iz * (s[4] - s[5]) + ix * iz * (s[7] - s[6]));
  dpz += scale * (ox * oy * (s[0] - s[4]) + ox * iy * (s[1] - s[5]) + ix * iy * (s[2] - s[6]) + ix * oy * (s[3] - s[7]));
}

// dL/d(normalised position p) -> dL/d(world position x); (x, y, z) is the un-contracted sample position
__device__ __forceinline__ void position_grad_to_world(const TnfModel& m, const float x, const float y, const float z,
                                                       const float sel, float& gx, float& gy, float& gz) {
  if (sel == 0.f) { gx = gy = gz = 0.f; return; }
  if (m.use_contraction) {
    gx *= 0.25f; gy *= 0.25f; gz *= 0.25f;
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    const float mag = fmaxf(ax, fmaxf(ay, az));
    if (!(mag < 1.f)) {
      // x' = a(m) x, a = (2 - 1/m) / m, m = |x|_inf = |x_k|:  J^T g = a g + (g . x) a'(m) sign(x_k) e_k
      const float a = (2.f - 1.f / mag) / mag;
      const float dadm = 2.f * (1.f - mag) / (mag * mag * mag);
      const float dot = gx * x + gy * y + gz * z;
      gx *= a; gy *= a; gz *= a;
      if (ax >= ay && ax >= az) gx += dot * dadm * (x < 0.f ? -1.f : 1.f);
      else if (ay >= az) gy += dot * dadm * (y < 0.f ? -1.f : 1.f);
      else gz += dot * dadm * (z < 0.f ? -1.f : 1.f);
    }
  } else {
    gx /= (m.aabb[3] - m.aabb[0]);
    gy /= (m.aabb[4] - m.aabb[1]);
    gz /= (m.aabb[5] - m.aabb[2]);
  }
}

// per-ray accumulation of the pose gradient: dL/do += sum_s g_s, dL/dd += sum_s mid_s g_s (warp-reduced)
__device__ __forceinline__ void flush_ray_grad(const TnfModelGrad& gr, const long long ray, float ax, float ay, float az,
                                               float bx, float by, float bz, const int lane) {
  ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
  bx = warp_sum(bx); by = warp_sum(by); bz = warp_sum(bz);
  if (lane == 0) {
    atomicAdd(gr.ray_origins + ray * 3 + 0, ax); atomicAdd(gr.ray_origins + ray * 3 + 1, ay);
    atomicAdd(gr.ray_origins + ray * 3 + 2, az);
    atomicAdd(gr.ray_directions + ray * 3 + 0, bx); atomicAdd(gr.ray_directions + ray * 3 + 1, by);
    atomicAdd(gr.ray_directions + ray * 3 + 2, bz);
  }
}

// reverse (suffix) exclusive scan helper over one 32-wide chunk: returns sum_{j>lane} v_j
__device__ __forceinline__ float warp_suffix_excl(float v, int lane, float& total) {
  const float rv = __shfl_sync(kFull, v, 31 - lane);
  const float incl = warp_incl_scan(rv, lane);
  total = __shfl_sync(kFull, incl, 31);
  return __shfl_sync(kFull, incl, 31 - lane) - v;
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf162(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void mma_16816_bf16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// ------------------------------------------------------------------------------------
// field level: shared per-ray prologue (compositing backward)
// ------------------------------------------------------------------------------------
struct FieldBwdScratch {
  float bins[kMaxFieldS + 8];
  float dsig[kMaxFieldS];  // dL/d sigma_i
  float dzr[kMaxFieldS], dzg[kMaxFieldS], dzb[kMaxFieldS];  // dL/d (pre-sigmoid colour)
  float dtau[kMaxFieldS];  // dL/d thermal_i
  float T[kMaxFieldS], w[kMaxFieldS], gw[kMaxFieldS];
  float rayb[64];
  float racc[64];  // TC kernel: per-ray column sums of dA1pre (kept here, not in 16 registers per lane)
};

// RGBRenderer / ThermalRenderer (background "last_sample"), AccumulationRenderer and get_weights,
// differentiated: fills dsig, dz*, dtau for the S2 samples of one ray.
__device__ __forceinline__ void composite_backward(const TnfModel& m, FieldBwdScratch& ws, const RayCtx& rc,
                                                   const int S2, const int lane, const float* __restrict__ fs,
                                                   const float* __restrict__ g_w2, const float gr_, const float gg_,
                                                   const float gb_, const float gth, const float gacc) {
  // pass 1: weights and transmittance from the saved densities
  float carry = 0.f, sw = 0.f;
  for (int base = 0; base < S2; base += 32) {
    const int i = base + lane;
    const bool active = i < S2;
    const int ii = active ? i : S2 - 1;
    float mid, delta;
    sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
    const float ds = active ? delta * fs[ii * 5] : 0.f;
    const float incl = warp_incl_scan(ds, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.f;
    const float T = expf(-(carry + excl));
    const float w = active ? (1.f - expf(-ds)) * T : 0.f;
    carry += __shfl_sync(kFull, incl, 31);
    if (active) { ws.T[i] = T; ws.w[i] = w; }
    sw += w;
  }
  sw = warp_sum(sw);
  // "last_sample" background (ThermalNerfModel) or none (concat_nerf: RGBTRenderer's "random" background)
  const bool concat = m.head_mode == TNF_HEAD_CONCAT;
  const float bgw = concat ? 0.f : 1.f - sw;
  const float* fl = fs + (S2 - 1) * 5;
  const float lr = concat ? 0.f : fl[1], lg = concat ? 0.f : fl[2], lb = concat ? 0.f : fl[3], lt = concat ? 0.f : fl[4];
  __syncwarp();
  // pass 2: dL/dw_i, then dL/d(delta*sigma)_i = gw_i (T_i - w_i) - sum_{j>i} gw_j w_j
  for (int i = lane; i < S2; i += 32) {
    const float* f = fs + i * 5;
    float g = gr_ * (f[1] - lr) + gg_ * (f[2] - lg) + gb_ * (f[3] - lb) + gth * (f[4] - lt) + gacc;
    if (g_w2) g += g_w2[i];
    ws.gw[i] = g;
    const float wc = ws.w[i] + (i == S2 - 1 ? bgw : 0.f);
    ws.dzr[i] = gr_ * wc * f[1] * (1.f - f[1]);
    ws.dzg[i] = gg_ * wc * f[2] * (1.f - f[2]);
    ws.dzb[i] = gb_ * wc * f[3] * (1.f - f[3]);
    // dL/d thermal_i (thermal head, linear output) or, for the RGBT head, dL/d(pre-sigmoid channel 3)
    ws.dtau[i] = concat ? gth * wc * f[4] * (1.f - f[4]) : gth * wc;
  }
  __syncwarp();
  float scarry = 0.f;
  for (int base = ((S2 - 1) / 32) * 32; base >= 0; base -= 32) {
    const int i = base + lane;
    const float v = i < S2 ? ws.gw[i] * ws.w[i] : 0.f;
    float tot;
    const float ex = warp_suffix_excl(v, lane, tot);
    if (i < S2) {
      float mid, delta;
      sample_geometry(rc, ws.bins[i], ws.bins[i + 1], mid, delta);
      ws.dsig[i] = (ws.gw[i] * (ws.T[i] - ws.w[i]) - (ex + scarry)) * delta;
    }
    scarry += tot;
  }
  __syncwarp();
}

// per-ray first-layer bias of the colour head, as in the forward, from global weights
__device__ __forceinline__ void ray_bias_and_inputs(const TnfModel& m, const TnfRays& rays, const long long ray,
                                                    const RayCtx& rc, const int lane, float* rayb, float (&sh)[16],
                                                    float& app_lane) {
  sh4((rc.dx + 1.f) * 0.5f, (rc.dy + 1.f) * 0.5f, (rc.dz + 1.f) * 0.5f, sh);
  const float* W = m.field.rgb0.weight;
  float e = 0.f;
  if (m.appearance_mode == TNF_APPEARANCE_LOOKUP) {
    e = __ldg(m.field.appearance + rays.camera_indices[ray] * 32 + lane);
  } else if (m.appearance_mode == TNF_APPEARANCE_MEAN) {
    for (int i = 0; i < m.field.num_images; ++i) e += m.field.appearance[i * 32 + lane];
    e /= (float)m.field.num_images;
  }
  app_lane = e;
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    const int n = lane + 32 * hlf;
    float b = m.field.rgb0.bias[n];
#pragma unroll
    for (int k = 0; k < 16; ++k) b = fmaf(sh[k], W[n * 63 + k], b);
    for (int j = 0; j < 32; ++j) b = fmaf(__shfl_sync(kFull, e, j), W[n * 63 + 31 + j], b);
    rayb[n] = b;
  }
}

}  // namespace tnf
