// Field / proposal MLP building blocks shared by the forward and backward kernels:
// shared-memory weight images, weight staging, the fp32 column-GEMV and the mma.sync tile
// helpers, per-ray geometry and the running compositor.
#pragma once

#include "tnf_device.cuh"

namespace tnf {

// ------------------------------------------------------------------------------------
// shared-memory images of the field MLPs
// ------------------------------------------------------------------------------------
struct FieldCommon {
  float rgb0sh_t[16 * 64];   // [k][n] SH block of mlp_head.layers.0 (cols 0..15)
  float rgb0app_t[32 * 64];  // [j][n] appearance block (cols 31..62)
  float rgb0b[64];           // bias (+ folded constant appearance part in eval)
  float app_const[32];
};

struct FieldW32 {
  float base0t[32 * 64];
  float base0b[64];
  float base1t[64 * 16];
  float base1b[16];
  float rgb0geo_t[16 * 64];  // row 0 = 0 (density slot), rows 1..15 = geo block (cols 16..30)
  float rgb1t[64 * 64];
  float rgb1b[64];
  float rgb2t[64 * 4];
  float rgb2b[4];
  float th0t[16 * 64];  // row 0 = 0
  float th0b[64];
  float th1t[64 * 64];
  float th1b[64];
  float th2[64];
  float th2b[4];
};

// mma.m16n8k16 B fragments, [k-tile][n-tile][lane]
struct FieldWTC {
  uint2 base0[2][8][32];
  uint2 base1[4][2][32];
  uint2 geo0[1][16][32];  // n-tiles 0..7: rgb0 geo block, 8..15: mlp_thermal.layers.0
  uint2 rgb1[4][8][32];
  uint2 rgb2[4][1][32];
  uint2 th1[4][8][32];
  uint2 th2[4][1][32];
  float base0b[64];
  float base1b[16];
  float rgb1b[64];
  float rgb2b[8];
  float th0b[64];
  float th1b[64];
  float th2b[8];
  float scal[TNF_MAX_LEVELS];  // grid.scalings: lanes index it with a per-lane level, which a constant-bank
                               // read would serialise (one pass per distinct address, long-scoreboard latency)
};

// ------------------------------------------------------------------------------------
// weight staging (once per CTA; weights are tiny and L2 resident)
// ------------------------------------------------------------------------------------
__device__ inline void stage_prop(PropW& W, const TnfDensityNet& net, int tid) {
  const int K = 2 * net.grid.num_levels;
  for (int i = tid; i < 256; i += kThreads) {
    const int k = i >> 4, j = i & 15;
    W.w0t[i] = (k < K) ? net.l0.weight[j * K + k] : 0.f;
  }
  if (tid < 16) {
    W.b0[tid] = net.l0.bias[tid];
    W.w1[tid] = net.l1.weight[tid];
  }
  if (tid == 0) W.b1 = net.l1.bias[0];
}

// logical [K][N] views (zero padded) of the torch [out,in] matrices
struct ViewPlain {  // W[k][n] = w[n*ld + k], k < kv, n < nv
  const float* w;
  int ld, kv, nv;
  __device__ float operator()(int k, int n) const { return (k < kv && n < nv) ? w[n * ld + k] : 0.f; }
};
struct ViewShift {  // row 0 is the (unused) density slot: W[k][n] = w[n*ld + col0 + k-1] for 1 <= k <= kv
  const float* w;
  int ld, col0, kv, nv;
  __device__ float operator()(int k, int n) const {
    return (k >= 1 && k <= kv && n < nv) ? w[n * ld + col0 + k - 1] : 0.f;
  }
};
struct ViewGeo0 {  // n < 64: rgb0 geo block, n >= 64: thermal layer 0
  ViewShift rgb, th;
  __device__ float operator()(int k, int n) const { return n < 64 ? rgb(k, n) : th(k, n - 64); }
};

template <typename V>
__device__ inline void stage_t(float* dst, int K, int N, const V& v, int tid) {
  for (int i = tid; i < K * N; i += kThreads) dst[i] = v(i / N, i % N);
}
// Position of the B fragment (k-tile kt, n-tile nt, lane) in a staged weight image with NT n-tiles per k-tile.
// For an even NT the fragments of n-tiles (2j, 2j+1) of one lane sit next to each other, so one 16-byte shared
// load feeds two mma instructions (half the LDS count of a fragment-per-load layout).
__device__ __forceinline__ int frag_index(int kt, int nt, int lane, int NT) {
  return (NT & 1) ? (kt * NT + nt) * 32 + lane : ((kt * (NT >> 1) + (nt >> 1)) * 32 + lane) * 2 + (nt & 1);
}
template <typename V>
__device__ inline void stage_frag(uint2* dst, int KT, int NT, const V& v, int tid, int nthreads = kThreads) {
  for (int i = tid; i < KT * NT * 32; i += nthreads) {
    const int lane = i & 31, nt = (i >> 5) % NT, kt = (i >> 5) / NT;
    const int g = lane >> 2, q = lane & 3;
    const int n = nt * 8 + g, k = kt * 16 + 2 * q;
    dst[frag_index(kt, nt, lane, NT)] =
        make_uint2(pack_half2(v(k, n), v(k + 1, n)), pack_half2(v(k + 8, n), v(k + 9, n)));
  }
}
__device__ inline void stage_vec(float* dst, const float* src, int n, int npad, int tid, int nthreads = kThreads) {
  for (int i = tid; i < npad; i += nthreads) dst[i] = i < n ? src[i] : 0.f;
}

__device__ inline void stage_common(FieldCommon& C, const TnfField& f, bool fold_appearance, int tid) {
  stage_t(C.rgb0sh_t, 16, 64, ViewPlain{f.rgb0.weight, 63, 16, 64}, tid);
  stage_t(C.rgb0app_t, 32, 64, ViewPlain{f.rgb0.weight + 31, 63, 32, 64}, tid);
  if (tid < 64) {
    float b = f.rgb0.bias[tid];
    if (fold_appearance) {
      for (int j = 0; j < 32; ++j) b = fmaf(C.app_const[j], f.rgb0.weight[tid * 63 + 31 + j], b);
    }
    C.rgb0b[tid] = b;
  }
}

__device__ inline void stage_field(FieldW32& W, const TnfField& f, int tid, int /*nthreads*/ = kThreads, int nout = 3) {
  stage_t(W.base0t, 32, 64, ViewPlain{f.base0.weight, 32, 32, 64}, tid);
  stage_t(W.base1t, 64, 16, ViewPlain{f.base1.weight, 64, 64, 16}, tid);
  stage_t(W.rgb0geo_t, 16, 64, ViewShift{f.rgb0.weight, 63, 16, 15, 64}, tid);
  stage_t(W.rgb1t, 64, 64, ViewPlain{f.rgb1.weight, 64, 64, 64}, tid);
  stage_t(W.rgb2t, 64, 4, ViewPlain{f.rgb2.weight, 64, 64, nout}, tid);  // nout = 4: RGBT head (concat_nerf)
  stage_t(W.th0t, 16, 64, ViewShift{f.th0.weight, 15, 0, 15, 64}, tid);
  stage_t(W.th1t, 64, 64, ViewPlain{f.th1.weight, 64, 64, 64}, tid);
  stage_vec(W.base0b, f.base0.bias, 64, 64, tid);
  stage_vec(W.base1b, f.base1.bias, 16, 16, tid);
  stage_vec(W.rgb1b, f.rgb1.bias, 64, 64, tid);
  stage_vec(W.rgb2b, f.rgb2.bias, nout, 4, tid);
  stage_vec(W.th0b, f.th0.bias, 64, 64, tid);
  stage_vec(W.th1b, f.th1.bias, 64, 64, tid);
  stage_vec(W.th2, f.th2.weight, 64, 64, tid);
  stage_vec(W.th2b, f.th2.bias, 1, 4, tid);
}

__device__ inline void stage_field(FieldWTC& W, const TnfField& f, int tid, int nthreads = kThreads, int nout = 3) {
  stage_frag(&W.base0[0][0][0], 2, 8, ViewPlain{f.base0.weight, 32, 32, 64}, tid, nthreads);
  stage_frag(&W.base1[0][0][0], 4, 2, ViewPlain{f.base1.weight, 64, 64, 16}, tid, nthreads);
  stage_frag(&W.geo0[0][0][0], 1, 16,
             ViewGeo0{ViewShift{f.rgb0.weight, 63, 16, 15, 64}, ViewShift{f.th0.weight, 15, 0, 15, 64}}, tid, nthreads);
  stage_frag(&W.rgb1[0][0][0], 4, 8, ViewPlain{f.rgb1.weight, 64, 64, 64}, tid, nthreads);
  stage_frag(&W.rgb2[0][0][0], 4, 1, ViewPlain{f.rgb2.weight, 64, 64, nout}, tid, nthreads);
  stage_frag(&W.th1[0][0][0], 4, 8, ViewPlain{f.th1.weight, 64, 64, 64}, tid, nthreads);
  stage_frag(&W.th2[0][0][0], 4, 1, ViewPlain{f.th2.weight, 64, 64, 1}, tid, nthreads);
  stage_vec(W.base0b, f.base0.bias, 64, 64, tid, nthreads);
  stage_vec(W.base1b, f.base1.bias, 16, 16, tid, nthreads);
  stage_vec(W.rgb1b, f.rgb1.bias, 64, 64, tid, nthreads);
  stage_vec(W.rgb2b, f.rgb2.bias, nout, 8, tid, nthreads);
  stage_vec(W.th0b, f.th0.bias, 64, 64, tid, nthreads);
  stage_vec(W.th1b, f.th1.bias, 64, 64, tid, nthreads);
  stage_vec(W.th2b, f.th2.bias, 1, 8, tid, nthreads);
  if (tid < TNF_MAX_LEVELS) W.scal[tid] = f.grid.scalings[tid];
}

// ------------------------------------------------------------------------------------
// per-ray state
// ------------------------------------------------------------------------------------
struct RayCtx {
  float ox, oy, oz, dx, dy, dz;
  float s_near, s_far;
};

// Cameras.generate_rays for one pixel of a perspective camera without distortion (nerfstudio 1.1.5
// Cameras._generate_rays_from_coords, reached from thermo_nerf/render/renderer.py:183): pixel centre
// (x+0.5, y+0.5), camera-frame direction ((x-cx)/fx, -(y-cy)/fy, -1), rotated by c2w[:3,:3] (products
// summed left to right as torch.sum over the last dim does for 3 elements), normalised; origin = c2w[:3,3].
__device__ __forceinline__ void camera_ray(const TnfCamera& c, const long long pix, float& ox, float& oy, float& oz,
                                           float& dx, float& dy, float& dz, float& norm) {
  const long long yi = pix / c.width;
  const float px = (float)(pix - yi * c.width) + 0.5f, py = (float)yi + 0.5f;
  const float a = (px - c.cx) / c.fx, b = -((py - c.cy) / c.fy), w = -1.f;
  const float x = __fadd_rn(__fadd_rn(__fmul_rn(a, c.c2w[0]), __fmul_rn(b, c.c2w[1])), __fmul_rn(w, c.c2w[2]));
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(a, c.c2w[4]), __fmul_rn(b, c.c2w[5])), __fmul_rn(w, c.c2w[6]));
  const float z = __fadd_rn(__fadd_rn(__fmul_rn(a, c.c2w[8]), __fmul_rn(b, c.c2w[9])), __fmul_rn(w, c.c2w[10]));
  norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  dx = x / norm;
  dy = y / norm;
  dz = z / norm;
  ox = c.c2w[3];
  oy = c.c2w[7];
  oz = c.c2w[11];
}

// frustums.get_positions(): origins + directions * (starts + ends) / 2, rounded like ATen (separate
// multiply and add, no FMA contraction) so that the discontinuous selector sees the same point.
__device__ __forceinline__ float ray_x(const RayCtx& rc, float mid) { return __fadd_rn(rc.ox, __fmul_rn(rc.dx, mid)); }
__device__ __forceinline__ float ray_y(const RayCtx& rc, float mid) { return __fadd_rn(rc.oy, __fmul_rn(rc.dy, mid)); }
__device__ __forceinline__ float ray_z(const RayCtx& rc, float mid) { return __fadd_rn(rc.oz, __fmul_rn(rc.dz, mid)); }

__device__ __forceinline__ void sample_geometry(const RayCtx& rc, float s0, float s1, float& mid, float& delta) {
  const float t0 = to_euclid(s0, rc.s_near, rc.s_far);
  const float t1 = to_euclid(s1, rc.s_near, rc.s_far);
  mid = (t0 + t1) * 0.5f;
  delta = t1 - t0;
}

// Running alpha-compositing state of one level (RaySamples.get_weights + median depth).
struct Compositor {
  float carry = 0.f;     // sum of delta*sigma of all previous samples
  float cw = 0.f;        // cumulative weight
  float median = 0.f;    // DepthRenderer("median")
  bool found = false;
  __device__ __forceinline__ float step(float ds, float mid, bool active, int lane) {
    const float incl = warp_incl_scan(ds, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.f;
    const float T = expf(-(carry + excl));
    const float alpha = 1.f - expf(-ds);
    const float w = active ? nan_to_num(alpha * T) : 0.f;
    carry += __shfl_sync(kFull, incl, 31);
    const float cwi = warp_incl_scan(w, lane) + cw;
    const unsigned hit = __ballot_sync(kFull, active && cwi >= 0.5f);
    if (!found && hit) {
      median = __shfl_sync(kFull, mid, __ffs(hit) - 1);
      found = true;
    }
    cw = __shfl_sync(kFull, cwi, 31);
    return w;
  }
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

template <int K, int N, int ACT>
__device__ __forceinline__ void dense_col(const float* __restrict__ wt, const float* __restrict__ bias,
                                          const float* xin, float (&y)[N]) {
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n);
    y[n] = b.x; y[n + 1] = b.y; y[n + 2] = b.z; y[n + 3] = b.w;
  }
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float x = xin[k * 32];
#pragma unroll
    for (int n = 0; n < N; n += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + k * N + n);
      y[n] = fmaf(x, w.x, y[n]);
      y[n + 1] = fmaf(x, w.y, y[n + 1]);
      y[n + 2] = fmaf(x, w.z, y[n + 2]);
      y[n + 3] = fmaf(x, w.w, y[n + 3]);
    }
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (ACT == ACT_RELU) y[n] = fmaxf(y[n], 0.f);
    if (ACT == ACT_SIGMOID) y[n] = sigmoidf(y[n]);
  }
}
template <int N>
__device__ __forceinline__ void store_col(float* xout, const float (&y)[N]) {
#pragma unroll
  for (int n = 0; n < N; ++n) xout[n * 32] = y[n];
}

template <int NT, int KT>
__device__ __forceinline__ void mma_layer(float (&c)[NT][4], const uint32_t (&a)[KT][4], const uint2* __restrict__ w,
                                          const int nt0, const int ntw, const int lane) {
  // w is a frag_index image with ntw n-tiles per k-tile; uses n-tiles nt0 .. nt0+NT-1 (nt0 even)
  if constexpr ((NT & 1) == 0) {
    const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(w);
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint4 b = w4[(kt * (ntw >> 1) + ((nt0 + nt) >> 1)) * 32 + lane];
        mma_16816(c[nt], a[kt], make_uint2(b.x, b.y));
        mma_16816(c[nt + 1], a[kt], make_uint2(b.z, b.w));
      }
  } else {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) mma_16816(c[nt], a[kt], w[frag_index(kt, nt0 + nt, lane, ntw)]);
  }
}
template <int NT>
__device__ __forceinline__ void init_bias(float (&c)[NT][4], const float* bias, const int q) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const float2 b = *reinterpret_cast<const float2*>(bias + nt * 8 + 2 * q);
    c[nt][0] = b.x; c[nt][1] = b.y; c[nt][2] = b.x; c[nt][3] = b.y;
  }
}
template <int NT, int ACT>
__device__ __forceinline__ void act_pack(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
  for (int kt = 0; kt < NT / 2; ++kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = c[2 * kt + h][e];
        if (ACT == ACT_RELU) v[e] = fmaxf(v[e], 0.f);
        if (ACT == ACT_SIGMOID) v[e] = sigmoid_fast(v[e]);
      }
      a[kt][2 * h] = pack_half2(v[0], v[1]);      // row g
      a[kt][2 * h + 1] = pack_half2(v[2], v[3]);  // row g+8
    }
  }
}

}  // namespace tnf
