// Fused forward of ThermalNerfModel.get_outputs
// (thermo_nerf/thermal_nerf/thermal_nerf_model.py:210-275) for sm_100a.
//
// One persistent kernel; one warp owns one ray from the first proposal sample to the
// composited pixel.  Nothing per-sample ever reaches HBM in eval mode:
//
//   level 0  256 piecewise-lindisp samples -> proposal net 0 (hash 5x2 -> 16 -> 1) -> weights
//   level 1  inverse-CDF resample (96)     -> proposal net 1                      -> weights
//   level 2  inverse-CDF resample (48)     -> field (hash 16x2 -> 64 -> 16 | rgb head | thermal head)
//   composite rgb / thermal (last-sample background), accumulation, median + expected depth
//
// Lanes map to consecutive samples of the ray so the 8 corner gathers of a level hit
// neighbouring (often identical) table cells -> few L1 wavefronts per load.  Weights, CDFs
// and bins live in three per-warp shared-memory lines that rotate between the levels; MLP
// weights are staged once per CTA.  TNF_PRECISION_TC_FP16 runs the 64-wide field MLPs on the
// tensor cores with register-resident activations (C fragments of one layer are the A
// fragments of the next); the hash features of a 16-sample tile go through a shared-memory
// tile (sample-major gathers, ldmatrix into A fragments, coalesced copy for the backward).
// Rays come from [R,3] tensors or are generated per pixel from a TnfCamera (Cameras.generate_rays).
// Also here: tnf_rays_kernel (stand-alone ray generation), tnf_post_kernel (Renderer.render's
// uint8 / colour-map conversion), tnf_clip_kernel (per-chunk expected-depth clip).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "tnf_field.cuh"

namespace tnf {

template <int PREC>
struct Smem;
template <>
struct Smem<TNF_PRECISION_FP32> {
  PropW prop[TNF_NUM_PROP];
  FieldCommon fc;
  FieldW32 fw;
  WarpScratch ws[kWarpsPerCta];
  float act[kWarpsPerCta][64 * 32];  // per-lane activation column: act[k*32 + lane]
  float geo[kWarpsPerCta][16 * 32];
};
constexpr int kFTileLd = 40;  // halves per row of the staged hash-feature tile: 80 B, conflict-free for ldmatrix
template <>
struct Smem<TNF_PRECISION_TC_FP16> {
  PropW prop[TNF_NUM_PROP];
  FieldCommon fc;
  FieldWTC fw;
  WarpScratch ws[kWarpsPerCta];
  __align__(16) __half ftile[kWarpsPerCta][16 * kFTileLd];  // hash features of one 16-sample tile, row-major
};

// ------------------------------------------------------------------------------------
// proposal level `lvl`: HashMLPDensityField.density_fn on the S samples between the spacing bins `bins`
// (S + 1 values in the warp's scratch) -> compositing weights in `wout`.  One instantiation serves both
// levels (runtime `lvl`): the kernel's instruction footprint has to stay inside the instruction cache.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float proposal_level(const TnfModel& m, const int lvl, const PropW& W,
                                                const float* __restrict__ bins, float* __restrict__ wout,
                                                const RayCtx& rc, const int S, const int lane,
                                                float* __restrict__ out_w, float* __restrict__ out_sdist) {
  const TnfDensityNet& net = m.prop[lvl];
  const int L = net.grid.num_levels;
  const uint32_t mask = (1u << net.grid.log2_size) - 1u;
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(net.grid.table);
  auto sb = [&](int i) -> float { return bins[i]; };

  Compositor comp;
  float last_mid = 0.f;
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool active = i < S;
    const int ii = active ? i : S - 1;
    float mid, delta;
    sample_geometry(rc, sb(ii), sb(ii + 1), mid, delta);
    float px, py, pz;
    const float sel = normalise_position(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), px, py, pz);
    float feat[2 * TNF_MAX_PROP_LEVELS];
#pragma unroll
    for (int l = 0; l < TNF_MAX_PROP_LEVELS; ++l) {
      float2 f = make_float2(0.f, 0.f);
      if (l < L) f = hash_level(tab + ((size_t)l << net.grid.log2_size), px, py, pz, net.grid.scalings[l], mask);
      feat[2 * l] = f.x;
      feat[2 * l + 1] = f.y;
    }
    float h[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(&W.b0[j]);
      h[j] = b.x; h[j + 1] = b.y; h[j + 2] = b.z; h[j + 3] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 2 * TNF_MAX_PROP_LEVELS; ++k) {
      if (k < 2 * L) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&W.w0t[k * 16 + j]);
          h[j] = fmaf(feat[k], w.x, h[j]);
          h[j + 1] = fmaf(feat[k], w.y, h[j + 1]);
          h[j + 2] = fmaf(feat[k], w.z, h[j + 2]);
          h[j + 3] = fmaf(feat[k], w.w, h[j + 3]);
        }
      }
    }
    float o = W.b1;
#pragma unroll
    for (int j = 0; j < 16; ++j) o = fmaf(fmaxf(h[j], 0.f), W.w1[j], o);
    const float density = expf(o) * sel;  // average_init_density == 1.0
    const float w = comp.step(active ? delta * density : 0.f, mid, active, lane);
    if (active) {
      wout[i] = w;
      if (out_w) out_w[i] = w;
    }
    if (base + 32 >= S) last_mid = __shfl_sync(kFull, mid, (S - 1) & 31);
  }
  if (out_sdist) {
    for (int i = lane; i <= S; i += 32) out_sdist[i] = sb(i);
  }
  __syncwarp();
  return comp.found ? comp.median : last_mid;
}

// ------------------------------------------------------------------------------------
// PDFSampler (include_original=False, histogram_padding=0.01, eps=1e-5): `w` (weights of the previous level,
// consumed) + `exist` (its Sprev + 1 spacing bins) -> `dst` (Snew + 1 spacing bins).  `dst` may alias `w`: every
// read of the weights happens before the first bin is written.
// ------------------------------------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ void pdf_resample(float* w_, float* __restrict__ cdf, const float* __restrict__ exist,
                                             float* dst, const int Sprev, const int Snew, const float anneal,
                                             const bool stratified, const float jit_new, const int lane) {
  float part = 0.f;
  for (int i = lane; i < Sprev; i += 32) {
    float w = w_[i];
    // tensor-core mode: w^a = exp2(a log2 w) on the SFU (w in [0,1]; 0 -> 0 for a > 0); fp32 mode: powf
    if (anneal != 1.f) w = FAST ? __powf(w, anneal) : powf(w, anneal);
    w += 0.01f;
    w_[i] = w;
    part += w;
  }
  float sum = warp_sum(part);
  const float padding = fmaxf(1e-5f - sum, 0.f);
  const float padw = padding / (float)Sprev;
  sum += padding;
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < Sprev; base += 32) {
    const int i = base + lane;
    const float p = (i < Sprev) ? (w_[i] + padw) / sum : 0.f;
    const float inc = warp_incl_scan(p, lane) + carry;
    if (i < Sprev) cdf[i + 1] = fminf(1.f, inc);
    carry = __shfl_sync(kFull, inc, 31);
  }
  __syncwarp();
  const int nb = Snew + 1;
  const float u_end = (float)(1.0 - 1.0 / (double)nb);
  const float u_off = stratified ? jit_new / (float)nb : (float)(1.0 / (double)(2 * nb));
  for (int j = lane; j < nb; j += 32) {
    const float u = linspace_at(j, nb, 0.f, u_end) + u_off;
    int lo = 0, hi = Sprev + 1;  // searchsorted(cdf, u, side="right")
    while (lo < hi) {
      const int midx = (lo + hi) >> 1;
      if (cdf[midx] <= u) lo = midx + 1; else hi = midx;
    }
    const int below = min(max(lo - 1, 0), Sprev);
    const int above = min(max(lo, 0), Sprev);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = exist[below], b1 = exist[above];
    float t = (u - c0) / (c1 - c0);
    t = isnan(t) ? 0.f : t;
    t = fminf(fmaxf(t, 0.f), 1.f);
    dst[j] = b0 + t * (b1 - b0);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------
// field level, fp32: lane per sample, activations in a per-lane shared-memory column
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void field_level(const TnfModel& m, Smem<TNF_PRECISION_FP32>& S, WarpScratch& ws,
                                            const RayCtx& rc, const int S2, const int lane, const int warp,
                                            void* __restrict__ save_feat) {
  const FieldW32& W = S.fw;
  float2* __restrict__ sf = reinterpret_cast<float2*>(save_feat);  // [S2][16] float2 of this ray, or null
  float* a = &S.act[warp][lane];
  float* geo = &S.geo[warp][lane];
  const TnfHashGrid& grid = m.field.grid;
  const uint32_t mask = (1u << grid.log2_size) - 1u;
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(grid.table);
  for (int base = 0; base < S2; base += 32) {
    const int i = base + lane;
    const bool active = i < S2;
    const int ii = active ? i : S2 - 1;
    float mid, delta;
    sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
    float px, py, pz;
    const float sel = normalise_position(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), px, py, pz);
#pragma unroll 4
    for (int l = 0; l < TNF_MAX_LEVELS; ++l) {
      const float2 f = hash_level(tab + ((size_t)l << grid.log2_size), px, py, pz, grid.scalings[l], mask);
      a[(2 * l) * 32] = f.x;
      a[(2 * l + 1) * 32] = f.y;
      if (sf && active) sf[i * 16 + l] = f;
    }
    {
      float y[64];
      dense_col<32, 64, ACT_RELU>(W.base0t, W.base0b, a, y);
      store_col(a, y);
    }
    float dba;
    {
      float y[16];
      dense_col<64, 16, ACT_NONE>(W.base1t, W.base1b, a, y);
      dba = y[0];
      y[0] = 0.f;  // density slot; its weight rows are zero
      store_col(geo, y);
    }
    float rgb[4];
    {
      float y[64];
      dense_col<16, 64, ACT_RELU>(W.rgb0geo_t, ws.rayb, geo, y);
      store_col(a, y);
      dense_col<64, 64, ACT_RELU>(W.rgb1t, W.rgb1b, a, y);
      store_col(a, y);
      dense_col<64, 4, ACT_SIGMOID>(W.rgb2t, W.rgb2b, a, rgb);
    }
    float th = rgb[3];  // concat_nerf: the fourth channel of the RGBT colour head (sigmoid applied)
    if (m.head_mode == TNF_HEAD_THERMAL) {
      float y[64];
      dense_col<16, 64, ACT_RELU>(W.th0t, W.th0b, geo, y);
      store_col(a, y);
      dense_col<64, 64, ACT_SIGMOID>(W.th1t, W.th1b, a, y);
      th = W.th2b[0];
#pragma unroll
      for (int k = 0; k < 64; ++k) th = fmaf(y[k], W.th2[k], th);
    }
    if (active) {
      ws.sigma[i] = expf(dba) * sel;
      ws.r[i] = rgb[0];
      ws.g[i] = rgb[1];
      ws.b[i] = rgb[2];
      ws.th[i] = th;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------
// field level, tensor cores: 16-sample tiles, mma.m16n8k16, activations stay in registers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void field_level(const TnfModel& m, Smem<TNF_PRECISION_TC_FP16>& S, WarpScratch& ws,
                                            const RayCtx& rc, const int S2, const int lane, const int warp,
                                            void* __restrict__ save_feat) {
  const FieldWTC& W = S.fw;
  uint32_t* __restrict__ sf = reinterpret_cast<uint32_t*>(save_feat);  // [S2][16] half2 of this ray, or null
  const TnfHashGrid& grid = m.field.grid;
  const uint32_t mask = (1u << grid.log2_size) - 1u;
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(grid.table);
  const int g = lane >> 2, q = lane & 3;
  __half* tile = S.ftile[warp];
  const int smp = lane & 15, par = lane >> 4;  // hash phase: this lane owns sample `smp`, levels of parity `par`
  for (int base = 0; base < S2; base += 16) {
    const int r0 = base + g, r1 = base + g + 8;
    // ---- hash encode, lanes = 16 consecutive samples x 2 level parities: one gather instruction covers two
    //      levels of 16 neighbouring samples, which at the coarser levels share grid cells (few sectors per
    //      request); a lane-per-(row, level-quad) mapping would touch 32 distinct lines every time.
    float selv;
    {
      const int ii = min(base + smp, S2 - 1);
      float mid, delta, px, py, pz;
      sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
      selv = normalise_position(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), px, py, pz);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int l = 2 * k + par;
        const float2 f = hash_level(tab + ((size_t)l << grid.log2_size), px, py, pz, W.scal[l], mask);
        *reinterpret_cast<uint32_t*>(tile + smp * kFTileLd + 2 * l) = pack_half2(f.x, f.y);
      }
    }
    __syncwarp();
    // ---- the tile as mma A fragments (two k-tiles of 16 halves) + the saved copy for the backward
    uint32_t a0[2][4];
    {
      const __half* src = tile + ((lane & 7) + ((lane >> 3) & 1) * 8) * kFTileLd + (lane >> 4) * 8;
      ldmatrix_x4(a0[0], src);
      ldmatrix_x4(a0[1], src + 16);
    }
    if (sf) {
#pragma unroll
      for (int c = lane; c < 64; c += 32) {
        const int row = c >> 2, part = c & 3;
        if (base + row < S2)
          reinterpret_cast<uint4*>(sf)[(base + row) * 4 + part] =
              *reinterpret_cast<const uint4*>(tile + row * kFTileLd + part * 8);
      }
    }
    const float sel[2] = {__shfl_sync(kFull, selv, g), __shfl_sync(kFull, selv, g + 8)};
    __syncwarp();  // every lane has read the tile before the next iteration overwrites it
    uint32_t hid[4][4];
    {
      float c[8][4];
      init_bias(c, W.base0b, q);
      mma_layer<8, 2>(c, a0, &W.base0[0][0][0], 0, 8, lane);
      act_pack<8, ACT_RELU>(c, hid);
    }
    float dba0, dba1;
    uint32_t ga[1][4];
    {
      float c[2][4];
      init_bias(c, W.base1b, q);
      mma_layer<2, 4>(c, hid, &W.base1[0][0][0], 0, 2, lane);
      dba0 = c[0][0];
      dba1 = c[0][2];
      if (q == 0) { c[0][0] = 0.f; c[0][2] = 0.f; }  // density slot (zero weight rows downstream)
      act_pack<2, ACT_NONE>(c, ga);
    }
    float rgb[1][4];
    {
      float c[8][4];
      init_bias(c, ws.rayb, q);
      mma_layer<8, 1>(c, ga, &W.geo0[0][0][0], 0, 16, lane);
      act_pack<8, ACT_RELU>(c, hid);
      init_bias(c, W.rgb1b, q);
      mma_layer<8, 4>(c, hid, &W.rgb1[0][0][0], 0, 8, lane);
      act_pack<8, ACT_RELU>(c, hid);
      init_bias(rgb, W.rgb2b, q);
      mma_layer<1, 4>(rgb, hid, &W.rgb2[0][0][0], 0, 1, lane);
    }
    float th[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    const bool concat = m.head_mode == TNF_HEAD_CONCAT;
    if (!concat) {
      float c[8][4];
      init_bias(c, W.th0b, q);
      mma_layer<8, 1>(c, ga, &W.geo0[0][0][0], 8, 16, lane);
      act_pack<8, ACT_RELU>(c, hid);
      init_bias(c, W.th1b, q);
      mma_layer<8, 4>(c, hid, &W.th1[0][0][0], 0, 8, lane);
      act_pack<8, ACT_SIGMOID>(c, hid);
      init_bias(th, W.th2b, q);
      mma_layer<1, 4>(th, hid, &W.th2[0][0][0], 0, 1, lane);
    }
    if (q == 0) {
      if (r0 < S2) {
        ws.sigma[r0] = expf(dba0) * sel[0];
        ws.r[r0] = sigmoid_fast(rgb[0][0]);
        ws.g[r0] = sigmoid_fast(rgb[0][1]);
        if (!concat) ws.th[r0] = th[0][0];
      }
      if (r1 < S2) {
        ws.sigma[r1] = expf(dba1) * sel[1];
        ws.r[r1] = sigmoid_fast(rgb[0][2]);
        ws.g[r1] = sigmoid_fast(rgb[0][3]);
        if (!concat) ws.th[r1] = th[0][2];
      }
    } else if (q == 1) {
      if (r0 < S2) ws.b[r0] = sigmoid_fast(rgb[0][0]);
      if (r1 < S2) ws.b[r1] = sigmoid_fast(rgb[0][2]);
      if (concat) {  // RGBT head: channel 3 is the temperature (rgb_concat/concat_field.py:65-75)
        if (r0 < S2) ws.th[r0] = sigmoid_fast(rgb[0][1]);
        if (r1 < S2) ws.th[r1] = sigmoid_fast(rgb[0][3]);
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------
// the kernel
//
// PHASE_ALL   the whole per-ray loop in one launch (eval / render: nothing per-sample reaches HBM)
// PHASE_PROP  the two proposal levels + both resamplings only; the 49 spacing bins of the final level go
//             to out.sdist[2] (training writes them anyway as ray_samples_list).  No tensor-core code in
//             this instantiation, so it compiles to half the registers and runs at twice the occupancy -
//             what the gather-latency-bound proposal levels want.
// PHASE_FIELD the field level + compositing, reading the bins PHASE_PROP wrote.
// ------------------------------------------------------------------------------------
enum { PHASE_ALL = 0, PHASE_PROP = 1, PHASE_FIELD = 2 };

struct SmemProp {
  PropW prop[TNF_NUM_PROP];
  WarpScratch ws[kWarpsPerCta];
};
template <int PREC, int PHASE>
struct SmemSel { using type = Smem<PREC>; };
template <int PREC>
struct SmemSel<PREC, PHASE_PROP> { using type = SmemProp; };

template <int PREC, int PHASE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
    tnf_forward_kernel(const __grid_constant__ TnfModel m, const __grid_constant__ TnfRays rays,
                       const __grid_constant__ TnfOutputs out, const long long chunk,
                       unsigned* __restrict__ clip_min, unsigned* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SmemT = typename SmemSel<PREC, PHASE>::type;
  SmemT& S = *reinterpret_cast<SmemT*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- stage weights
  const bool lookup = m.appearance_mode == TNF_APPEARANCE_LOOKUP;
  if constexpr (PHASE != PHASE_FIELD) {
    stage_prop(S.prop[0], m.prop[0], tid);
    stage_prop(S.prop[1], m.prop[1], tid);
  }
  if constexpr (PHASE != PHASE_PROP) {
    if (warp == 0) {
      float e = 0.f;
      if (m.appearance_mode == TNF_APPEARANCE_MEAN) {
        for (int i = 0; i < m.field.num_images; ++i) e += m.field.appearance[i * 32 + lane];
        e /= (float)m.field.num_images;
      }
      S.fc.app_const[lane] = e;
    }
    stage_field(S.fw, m.field, tid, kThreads, m.head_mode == TNF_HEAD_CONCAT ? 4 : 3);
    __syncthreads();
    stage_common(S.fc, m.field, !lookup, tid);
  }
  __syncthreads();

  WarpScratch& ws = S.ws[warp];
  const bool train = m.training != 0;
  const bool stratified = train && rays.jitter != nullptr;
  const int S0 = m.num_samples[0], S2 = m.num_samples[2];
  const long long R = rays.num_rays;
  long long cur_chunk = -1;
  float cmin = FLT_MAX, cmax = 0.f;

  for (long long ray = (long long)blockIdx.x * kWarpsPerCta + warp; ray < R;
       ray += (long long)gridDim.x * kWarpsPerCta) {
    RayCtx rc;
    if (rays.from_camera) {
      float nrm;
      camera_ray(rays.camera, rays.first_pixel + ray, rc.ox, rc.oy, rc.oz, rc.dx, rc.dy, rc.dz, nrm);
    } else {
      rc.ox = __ldg(rays.origins + ray * 3 + 0);
      rc.oy = __ldg(rays.origins + ray * 3 + 1);
      rc.oz = __ldg(rays.origins + ray * 3 + 2);
      rc.dx = __ldg(rays.directions + ray * 3 + 0);
      rc.dy = __ldg(rays.directions + ray * 3 + 1);
      rc.dz = __ldg(rays.directions + ray * 3 + 2);
    }
    const float near = rays.nears ? __ldg(rays.nears + ray) : m.near_plane;
    const float far = rays.fars ? __ldg(rays.fars + ray) : m.far_plane;
    rc.s_near = spacing_fn(near);
    rc.s_far = spacing_fn(far);
    float jit0 = 0.f, jit1 = 0.f, jit2 = 0.f;
    if (stratified) {
      jit0 = __ldg(rays.jitter + ray);
      jit1 = __ldg(rays.jitter + R + ray);
      jit2 = __ldg(rays.jitter + 2 * R + ray);
    }

    // ---- per-ray first-layer bias of the colour head: bias + W_sh * SH((d+1)/2) [+ W_app * e_cam]
    if constexpr (PHASE != PHASE_PROP) {
      float sh[16];
      sh4((rc.dx + 1.f) * 0.5f, (rc.dy + 1.f) * 0.5f, (rc.dz + 1.f) * 0.5f, sh);
      float b0 = S.fc.rgb0b[lane], b1 = S.fc.rgb0b[lane + 32];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        b0 = fmaf(sh[k], S.fc.rgb0sh_t[k * 64 + lane], b0);
        b1 = fmaf(sh[k], S.fc.rgb0sh_t[k * 64 + lane + 32], b1);
      }
      if (lookup) {
        const long long cam = rays.camera_indices[ray];
        const float e = __ldg(m.field.appearance + cam * 32 + lane);
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const float ej = __shfl_sync(kFull, e, j);
          b0 = fmaf(ej, S.fc.rgb0app_t[j * 64 + lane], b0);
          b1 = fmaf(ej, S.fc.rgb0app_t[j * 64 + lane + 32], b1);
        }
      }
      ws.rayb[lane] = b0;
      ws.rayb[lane + 32] = b1;
    }

    float pd0 = 0.f, pd1 = 0.f;
    if constexpr (PHASE != PHASE_FIELD) {
      // three scratch lines rotate: A = bins of the current level, B = its weights -> the next level's bins,
      // C = cdf.  After the two levels the final level's bins sit in ws.bins again.
      float* A = ws.bins;
      float* B = ws.w;
      for (int i = lane; i <= S0; i += 32) A[i] = initial_sbin(i, S0, stratified, jit0);
      __syncwarp();
#pragma unroll 1
      for (int lvl = 0; lvl < TNF_NUM_PROP; ++lvl) {
        const int Sc = m.num_samples[lvl], Sn = m.num_samples[lvl + 1];
        const float pd = proposal_level(m, lvl, S.prop[lvl], A, B, rc, Sc, lane,
                                        out.weights[lvl] ? out.weights[lvl] + ray * Sc : nullptr,
                                        out.sdist[lvl] ? out.sdist[lvl] + ray * (Sc + 1) : nullptr);
        if (lvl == 0) pd0 = pd; else pd1 = pd;
        pdf_resample<PREC == TNF_PRECISION_TC_FP16>(B, ws.cdf, A, B, Sc, Sn, m.anneal, stratified,
                                                    lvl == 0 ? jit1 : jit2, lane);
        float* t = A; A = B; B = t;
      }
    }
    if constexpr (PHASE == PHASE_PROP) {
      // hand the final level's spacing bins (and the two proposal depths) to the field launch
      for (int i = lane; i <= S2; i += 32) out.sdist[2][ray * (S2 + 1) + i] = ws.bins[i];
      if (lane == 0) {
        out.prop_depth[0][ray] = pd0;
        out.prop_depth[1][ray] = pd1;
      }
      __syncwarp();
      continue;
    }
    if constexpr (PHASE == PHASE_FIELD) {
      for (int i = lane; i <= S2; i += 32) ws.bins[i] = out.sdist[2][ray * (S2 + 1) + i];
      __syncwarp();
    }
    if constexpr (PHASE != PHASE_PROP) {
    // ---- level 2: field; its spacing bins live in ws.bins[0..S2]
    field_level(m, S, ws, rc, S2, lane, warp,
                out.field_features
                    ? static_cast<unsigned char*>(out.field_features) +
                          (size_t)ray * S2 * 32 * (PREC == TNF_PRECISION_TC_FP16 ? 2 : 4)
                    : nullptr);

    // ---- composite (get_weights + RGB/Thermal/Accumulation/Depth renderers)
    Compositor comp;
    float sr = 0.f, sg = 0.f, sb = 0.f, st = 0.f, sw = 0.f, swt = 0.f;
    float first_mid = 0.f, last_mid = 0.f;
    for (int base = 0; base < S2; base += 32) {
      const int i = base + lane;
      const bool active = i < S2;
      const int ii = active ? i : S2 - 1;
      float mid, delta;
      sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
      const float w = comp.step(active ? delta * ws.sigma[ii] : 0.f, mid, active, lane);
      float cr = ws.r[ii], cg = ws.g[ii], cb = ws.b[ii], ct = ws.th[ii];
      if (!train) { cr = nan_to_num(cr); cg = nan_to_num(cg); cb = nan_to_num(cb); ct = nan_to_num(ct); }
      sr = fmaf(w, cr, sr);
      sg = fmaf(w, cg, sg);
      sb = fmaf(w, cb, sb);
      st = fmaf(w, ct, st);
      sw += w;
      swt = fmaf(w, mid, swt);
      if (active && out.weights[2]) out.weights[2][ray * S2 + i] = w;
      if (active && out.field_samples) {
        float* fs = out.field_samples + ((size_t)ray * S2 + i) * 5;
        fs[0] = ws.sigma[ii]; fs[1] = cr; fs[2] = cg; fs[3] = cb; fs[4] = ct;
      }
      if (base == 0) first_mid = __shfl_sync(kFull, mid, 0);
      if (base + 32 >= S2) last_mid = __shfl_sync(kFull, mid, (S2 - 1) & 31);
    }
    if (PHASE == PHASE_ALL && out.sdist[2]) {
      for (int i = lane; i <= S2; i += 32) out.sdist[2][ray * (S2 + 1) + i] = ws.bins[i];
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); st = warp_sum(st);
    sw = warp_sum(sw); swt = warp_sum(swt);
    if (lane == 0) {
      float lr = ws.r[S2 - 1], lg = ws.g[S2 - 1], lb = ws.b[S2 - 1], lt = ws.th[S2 - 1];
      if (!train) { lr = nan_to_num(lr); lg = nan_to_num(lg); lb = nan_to_num(lb); lt = nan_to_num(lt); }
      // background_color = "last_sample" (thermal_renderer.py:49); concat_nerf: RGBTRenderer's default "random"
      // background adds nothing to the composite (rgbt_renderer.py:63-71)
      const float bgw = m.head_mode == TNF_HEAD_CONCAT ? 0.f : 1.f - sw;
      float r = sr + lr * bgw, g = sg + lg * bgw, b = sb + lb * bgw, t = st + lt * bgw;
      if (!train) {
        r = fminf(fmaxf(r, 0.f), 1.f); g = fminf(fmaxf(g, 0.f), 1.f);
        b = fminf(fmaxf(b, 0.f), 1.f); t = fminf(fmaxf(t, 0.f), 1.f);
      }
      out.rgb[ray * 3 + 0] = r;
      out.rgb[ray * 3 + 1] = g;
      out.rgb[ray * 3 + 2] = b;
      out.thermal[ray] = t;
      out.accumulation[ray] = sw;
      out.depth[ray] = comp.found ? comp.median : last_mid;
      out.expected_depth[ray] = swt / (sw + 1e-10f);  // clipped by tnf_clip_kernel
      if (PHASE == PHASE_ALL) {
        out.prop_depth[0][ray] = pd0;
        out.prop_depth[1][ray] = pd1;
      }
    }
    // ---- per-chunk min/max of the sample mid-points (tensor-global clip of DepthRenderer("expected"))
    const long long c = chunk > 0 ? ray / chunk : 0;
    if (c != cur_chunk) {
      if (cur_chunk >= 0 && lane == 0) {
        atomicMin(clip_min + cur_chunk, __float_as_uint(cmin));
        atomicMax(clip_max + cur_chunk, __float_as_uint(cmax));
      }
      cur_chunk = c;
      cmin = FLT_MAX;
      cmax = 0.f;
    }
    if (first_mid == first_mid) cmin = fminf(cmin, first_mid);
    if (last_mid == last_mid) cmax = fmaxf(cmax, last_mid);
    __syncwarp();
    }  // PHASE != PHASE_PROP
  }
  if (PHASE != PHASE_PROP && cur_chunk >= 0 && lane == 0) {
    atomicMin(clip_min + cur_chunk, __float_as_uint(cmin));
    atomicMax(clip_max + cur_chunk, __float_as_uint(cmax));
  }
}

__global__ void tnf_clip_kernel(float* __restrict__ expected_depth, const long long R, const long long chunk,
                                const unsigned* __restrict__ clip_min, const unsigned* __restrict__ clip_max) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const long long c = chunk > 0 ? r / chunk : 0;
  const float lo = __uint_as_float(clip_min[c]), hi = __uint_as_float(clip_max[c]);
  const float e = expected_depth[r];
  if (e == e) expected_depth[r] = fminf(fmaxf(e, lo), hi);
}

// Cameras.generate_rays as a stand-alone pass (callers that want the RayBundle tensors themselves).
__global__ void tnf_rays_kernel(const __grid_constant__ TnfCamera cam, const long long first, const long long n,
                                float* __restrict__ origins, float* __restrict__ directions,
                                float* __restrict__ dnorm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ox, oy, oz, dx, dy, dz, nrm;
  camera_ray(cam, first + i, ox, oy, oz, dx, dy, dz, nrm);
  origins[3 * i] = ox; origins[3 * i + 1] = oy; origins[3 * i + 2] = oz;
  directions[3 * i] = dx; directions[3 * i + 1] = dy; directions[3 * i + 2] = dz;
  if (dnorm) dnorm[i] = nrm;
}

// Renderer.render's per-frame conversion (thermo_nerf/render/renderer.py:189-199): float image -> uint8,
// scalar image -> colour-mapped (or grey) uint8; the colour table sits in shared memory.
__device__ __forceinline__ unsigned char to_u8(float v) {  // numpy (x * 255).astype(uint8) for x in [0,1]
  return (unsigned char)(int)__fmul_rn(v, 255.f);
}
__global__ void tnf_post_kernel(const float* __restrict__ rgb, const float* __restrict__ scalar, const long long n,
                                const unsigned char* __restrict__ lut8, const int lut_n,
                                unsigned char* __restrict__ rgb8, unsigned char* __restrict__ scalar8) {
  extern __shared__ unsigned char s_lut[];
  if (lut8)
    for (int i = threadIdx.x; i < 3 * lut_n; i += blockDim.x) s_lut[i] = lut8[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (rgb && rgb8) {
      rgb8[3 * i] = to_u8(rgb[3 * i]);
      rgb8[3 * i + 1] = to_u8(rgb[3 * i + 1]);
      rgb8[3 * i + 2] = to_u8(rgb[3 * i + 2]);
    }
    if (scalar && scalar8) {
      const float v = scalar[i];
      unsigned char r, g, b;
      if (lut8) {
        if (v != v) {
          r = g = b = 0;  // matplotlib's "bad" colour is transparent black
        } else {
          const float xa = __fmul_rn(v, (float)lut_n);
          int k = xa == (float)lut_n ? lut_n - 1 : (int)xa;
          k = xa < 0.f ? 0 : min(k, lut_n - 1);  // default under / over colours are the end colours
          r = s_lut[3 * k]; g = s_lut[3 * k + 1]; b = s_lut[3 * k + 2];
        }
      } else {
        r = g = b = to_u8(v);
      }
      scalar8[3 * i] = r; scalar8[3 * i + 1] = g; scalar8[3 * i + 2] = b;
    }
  }
}

}  // namespace tnf

// ====================================================================================
// C ABI
// ====================================================================================
#include "tnf_host.h"

namespace {
using tnf::fail;
using tnf::g_err;
using tnf::aligned16;

int check_linear(const TnfLinear& l, const char* name) {
  if (!l.weight || !l.bias) return fail(TNF_ERR_INVALID_ARGUMENT, "%s: null weight/bias", name);
  return TNF_OK;
}
}  // namespace

namespace tnf {
// The run-merged scatter identifies a grid cell by its floor coordinates packed into 11 bits each (HashCorners::k1):
// unique for grid resolutions below 2048.  nerfacto's default field tops out at 2047; nerfacto-big / -huge
// (max_res 4096 / 8192) would alias cells and corrupt hash-table gradients, so they are rejected here.
static int check_scalings(const TnfHashGrid& g, const char* name) {
  for (int l = 0; l < g.num_levels; ++l)
    if (!(g.scalings[l] > 0.f) || !(g.scalings[l] < 2048.f))
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "%s: scalings[%d]=%g outside (0, 2048): the kernels are built for "
                  "max_res < 2048", name, l, (double)g.scalings[l]);
  return TNF_OK;
}

int check_model(const TnfModel* m) {
  if (!m) return fail(TNF_ERR_INVALID_ARGUMENT, "model is null");
  for (int k = 0; k < TNF_NUM_PROP; ++k) {
    const TnfDensityNet& n = m->prop[k];
    if (!n.grid.table) return fail(TNF_ERR_INVALID_ARGUMENT, "prop[%d].grid.table is null", k);
    if (n.grid.num_levels < 1 || n.grid.num_levels > TNF_MAX_PROP_LEVELS)
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "prop[%d]: num_levels=%d not in [1,%d]", k, n.grid.num_levels,
                  TNF_MAX_PROP_LEVELS);
    if (n.grid.log2_size < 1 || n.grid.log2_size > 24)
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "prop[%d]: log2_size=%d not in [1,24]", k, n.grid.log2_size);
    if (int e = check_linear(n.l0, "prop.l0")) return e;
    if (int e = check_linear(n.l1, "prop.l1")) return e;
    if (int e = check_scalings(n.grid, "prop.grid")) return e;
  }
  const TnfField& f = m->field;
  if (!f.grid.table) return fail(TNF_ERR_INVALID_ARGUMENT, "field.grid.table is null");
  if (f.grid.num_levels != TNF_MAX_LEVELS)
    return fail(TNF_ERR_UNSUPPORTED_CONFIG, "field: num_levels=%d, kernels are built for %d", f.grid.num_levels,
                TNF_MAX_LEVELS);
  if (f.grid.log2_size < 1 || f.grid.log2_size > 24)
    return fail(TNF_ERR_UNSUPPORTED_CONFIG, "field: log2_size=%d not in [1,24]", f.grid.log2_size);
  if (int e = check_scalings(f.grid, "field.grid")) return e;
  const TnfLinear* ls[] = {&f.base0, &f.base1, &f.rgb0, &f.rgb1, &f.rgb2, &f.th0, &f.th1, &f.th2};
  for (const TnfLinear* l : ls)
    if (int e = check_linear(*l, "field linear")) return e;
  if (m->appearance_mode != TNF_APPEARANCE_ZEROS && (!f.appearance || f.num_images < 1))
    return fail(TNF_ERR_INVALID_ARGUMENT, "field.appearance is required for appearance_mode=%d",
                m->appearance_mode);
  if (m->appearance_mode < 0 || m->appearance_mode > 2)
    return fail(TNF_ERR_INVALID_ARGUMENT, "appearance_mode=%d", m->appearance_mode);
  for (int k = 0; k <= TNF_NUM_PROP; ++k) {
    const int lim = (k == TNF_NUM_PROP) ? tnf::kMaxFieldS : TNF_MAX_SAMPLES;
    if (m->num_samples[k] < 1 || m->num_samples[k] > lim)
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "num_samples[%d]=%d not in [1,%d]", k, m->num_samples[k], lim);
  }
  if (m->precision != TNF_PRECISION_FP32 && m->precision != TNF_PRECISION_TC_FP16)
    return fail(TNF_ERR_INVALID_ARGUMENT, "precision=%d", m->precision);
  if (m->head_mode != TNF_HEAD_THERMAL && m->head_mode != TNF_HEAD_CONCAT)
    return fail(TNF_ERR_INVALID_ARGUMENT, "head_mode=%d", m->head_mode);
  return TNF_OK;
}
}  // namespace tnf

namespace {
using tnf::check_model;

template <int PREC, int PHASE, int MINB>
int launch_forward(const TnfModel& m, const TnfRays& r, const TnfOutputs& o, long long chunk, unsigned* cmin,
                   unsigned* cmax, cudaStream_t stream) {
  static thread_local int configured_dev = -1;
  const int num_sms = tnf::num_sms();
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  const size_t smem = sizeof(typename tnf::SmemSel<PREC, PHASE>::type);
  if (configured_dev != dev) {
    e = cudaFuncSetAttribute(tnf::tnf_forward_kernel<PREC, PHASE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    if (e != cudaSuccess)
      return fail(TNF_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    configured_dev = dev;
  }
  long long want = (r.num_rays + tnf::kWarpsPerCta - 1) / tnf::kWarpsPerCta;
  const long long cap = (long long)num_sms * MINB;
  const int grid = (int)(want < cap ? want : cap);
  tnf::tnf_forward_kernel<PREC, PHASE, MINB><<<grid, tnf::kThreads, smem, stream>>>(m, r, o, chunk, cmin, cmax);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "forward kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

// Training forward in tensor-core mode (needs out.sdist[2]).  0: one fused launch; 3 / 4: proposal launch at that many
// CTAs per SM + field launch.  Measured on B200 at 4096 rays (profiles/r2_forward_split.json): 4 -> 0.212 ms,
// 0 -> 0.221 ms, 3 -> 0.229 ms (5 CTAs per SM = 48 registers, 136 B of spills: 0.236 ms); TNF_FORWARD_SPLIT overrides.
int split_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* v = getenv("TNF_FORWARD_SPLIT");
    mode = v ? atoi(v) : 4;
    if (mode != 0 && mode != 3 && mode != 4) mode = 4;
  }
  return mode;
}
}  // namespace

extern "C" {

int tnf_version(void) { return TNF_ABI_VERSION; }

const char* tnf_last_error(void) { return g_err; }

size_t tnf_forward_workspace_bytes(int64_t num_rays, int64_t depth_clip_chunk) {
  if (num_rays <= 0) return 16;
  const int64_t chunks = depth_clip_chunk > 0 ? (num_rays + depth_clip_chunk - 1) / depth_clip_chunk : 1;
  return (size_t)(2 * chunks * sizeof(unsigned) + 16);
}

int tnf_render_forward(const TnfModel* model, const TnfRays* rays, const TnfOutputs* out, int64_t depth_clip_chunk,
                       void* workspace, size_t workspace_bytes, void* stream_) {
  return tnf_render_forward_staged(model, rays, out, depth_clip_chunk, workspace, workspace_bytes, stream_, nullptr);
}

int tnf_render_forward_staged(const TnfModel* model, const TnfRays* rays, const TnfOutputs* out,
                              int64_t depth_clip_chunk, void* workspace, size_t workspace_bytes, void* stream_,
                              void* field_params_ready) {
  g_err[0] = 0;
  if (int e = check_model(model)) return e;
  if (!rays || !out) return fail(TNF_ERR_INVALID_ARGUMENT, "rays/out is null");
  if (rays->num_rays < 0) return fail(TNF_ERR_INVALID_ARGUMENT, "num_rays=%lld", (long long)rays->num_rays);
  if (rays->num_rays == 0) return TNF_OK;
  if (rays->from_camera) {
    const TnfCamera& c = rays->camera;
    if (c.width < 1 || c.height < 1 || !(c.fx != 0.f) || !(c.fy != 0.f))
      return fail(TNF_ERR_INVALID_ARGUMENT, "camera: width=%d height=%d fx=%g fy=%g", c.width, c.height, c.fx, c.fy);
    if (rays->first_pixel < 0 || rays->first_pixel + rays->num_rays > (int64_t)c.width * c.height)
      return fail(TNF_ERR_INVALID_ARGUMENT, "camera: pixels [%lld, %lld) outside the %dx%d image",
                  (long long)rays->first_pixel, (long long)(rays->first_pixel + rays->num_rays), c.width, c.height);
    if (model->training) return fail(TNF_ERR_INVALID_ARGUMENT, "from_camera rays are an eval-mode input");
  } else if (!rays->origins || !rays->directions) {
    return fail(TNF_ERR_INVALID_ARGUMENT, "origins/directions is null");
  }
  if (model->appearance_mode == TNF_APPEARANCE_LOOKUP && !rays->camera_indices)
    return fail(TNF_ERR_INVALID_ARGUMENT, "camera_indices required for TNF_APPEARANCE_LOOKUP");
  if (!out->rgb || !out->thermal || !out->depth || !out->expected_depth || !out->accumulation ||
      !out->prop_depth[0] || !out->prop_depth[1])
    return fail(TNF_ERR_INVALID_ARGUMENT, "a required output pointer is null");
  if ((!rays->from_camera && (!aligned16(rays->origins) || !aligned16(rays->directions))) || !aligned16(out->rgb))
    return fail(TNF_ERR_INVALID_ARGUMENT, "ray/output buffers must be 16-byte aligned");
  const size_t need = tnf_forward_workspace_bytes(rays->num_rays, depth_clip_chunk);
  if (!workspace || workspace_bytes < need)
    return fail(TNF_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu bytes", workspace_bytes, need);

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long chunk = depth_clip_chunk > 0 ? depth_clip_chunk : 0;
  const long long chunks = chunk > 0 ? (rays->num_rays + chunk - 1) / chunk : 1;
  unsigned* cmin = static_cast<unsigned*>(workspace);
  unsigned* cmax = cmin + chunks;
  cudaError_t e = cudaMemsetAsync(cmin, 0x7f, chunks * sizeof(unsigned), stream);  // 3.39e38: above any depth
  if (e == cudaSuccess) e = cudaMemsetAsync(cmax, 0, chunks * sizeof(unsigned), stream);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));

  int rc;
  using namespace tnf;
  int split = (out->sdist[2] && model->precision == TNF_PRECISION_TC_FP16) ? split_mode() : 0;
  // an exchange to hide: take the two-launch form even where the fused launch was asked for
  if (field_params_ready && split == 0 && out->sdist[2] && model->precision == TNF_PRECISION_TC_FP16) split = 4;
  auto wait_field = [&]() -> int {
    if (!field_params_ready) return TNF_OK;
    const cudaError_t we = cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(field_params_ready), 0);
    return we == cudaSuccess ? TNF_OK : fail(TNF_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(we));
  };
  if (split == 0)
    if (int wrc = wait_field()) return wrc;
  if (split == 4) {
    rc = launch_forward<TNF_PRECISION_TC_FP16, PHASE_PROP, 4>(*model, *rays, *out, chunk, cmin, cmax, stream);
    if (rc == TNF_OK) rc = wait_field();
    if (rc == TNF_OK) rc = launch_forward<TNF_PRECISION_TC_FP16, PHASE_FIELD, 2>(*model, *rays, *out, chunk, cmin, cmax, stream);
  } else if (split == 3) {
    rc = launch_forward<TNF_PRECISION_TC_FP16, PHASE_PROP, 3>(*model, *rays, *out, chunk, cmin, cmax, stream);
    if (rc == TNF_OK) rc = wait_field();
    if (rc == TNF_OK) rc = launch_forward<TNF_PRECISION_TC_FP16, PHASE_FIELD, 2>(*model, *rays, *out, chunk, cmin, cmax, stream);
  } else if (model->precision == TNF_PRECISION_TC_FP16) {
    rc = launch_forward<TNF_PRECISION_TC_FP16, PHASE_ALL, 2>(*model, *rays, *out, chunk, cmin, cmax, stream);
  } else {
    rc = launch_forward<TNF_PRECISION_FP32, PHASE_ALL, 1>(*model, *rays, *out, chunk, cmin, cmax, stream);
  }
  if (rc != TNF_OK) return rc;

  const int tb = 256;
  const long long nb = (rays->num_rays + tb - 1) / tb;
  tnf::tnf_clip_kernel<<<(unsigned)nb, tb, 0, stream>>>(out->expected_depth, rays->num_rays, chunk, cmin, cmax);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "clip kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

int tnf_generate_rays(const TnfCamera* camera, int64_t first_pixel, int64_t num_pixels, float* origins,
                      float* directions, float* directions_norm, void* stream_) {
  g_err[0] = 0;
  if (!camera) return fail(TNF_ERR_INVALID_ARGUMENT, "camera is null");
  if (camera->width < 1 || camera->height < 1 || !(camera->fx != 0.f) || !(camera->fy != 0.f))
    return fail(TNF_ERR_INVALID_ARGUMENT, "camera: width=%d height=%d fx=%g fy=%g", camera->width, camera->height,
                camera->fx, camera->fy);
  if (num_pixels < 0 || first_pixel < 0 || first_pixel + num_pixels > (int64_t)camera->width * camera->height)
    return fail(TNF_ERR_INVALID_ARGUMENT, "pixels [%lld, %lld) outside the %dx%d image", (long long)first_pixel,
                (long long)(first_pixel + num_pixels), camera->width, camera->height);
  if (num_pixels == 0) return TNF_OK;
  if (!origins || !directions) return fail(TNF_ERR_INVALID_ARGUMENT, "origins/directions is null");
  const int tb = 256;
  tnf::tnf_rays_kernel<<<(unsigned)((num_pixels + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream_)>>>(
      *camera, first_pixel, num_pixels, origins, directions, directions_norm);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "rays kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

int tnf_postprocess_frame(const float* rgb, const float* scalar, int64_t num_pixels, const uint8_t* lut8,
                          int32_t lut_n, uint8_t* rgb8, uint8_t* scalar8, void* stream_) {
  g_err[0] = 0;
  if (num_pixels < 0) return fail(TNF_ERR_INVALID_ARGUMENT, "num_pixels=%lld", (long long)num_pixels);
  if ((rgb == nullptr) != (rgb8 == nullptr) || (scalar == nullptr) != (scalar8 == nullptr))
    return fail(TNF_ERR_INVALID_ARGUMENT, "an input image and its uint8 output must be given together");
  if (lut8 && (lut_n < 1 || lut_n > 4096)) return fail(TNF_ERR_INVALID_ARGUMENT, "lut_n=%d not in [1,4096]", lut_n);
  if (num_pixels == 0 || (!rgb && !scalar)) return TNF_OK;
  const int tb = 256;
  long long nb = (num_pixels + tb - 1) / tb;
  const long long cap = (long long)tnf::num_sms() * 8;
  if (nb > cap) nb = cap;
  tnf::tnf_post_kernel<<<(unsigned)nb, tb, lut8 ? 3 * (size_t)lut_n : 0, static_cast<cudaStream_t>(stream_)>>>(
      rgb, scalar, num_pixels, lut8, lut_n, rgb8, scalar8);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "post kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

}  // extern "C"
