// Backward of ThermalNerfModel.get_outputs (thermo_nerf/thermal_nerf/thermal_nerf_model.py:210-275)
// as driven by get_loss_dict (:277-326), for sm_100a.
//
//   tnf_backward_prop_kernel   the two HashMLPDensityField proposal levels (:127-148) as (level, ray) units from
//                              a device counter: re-gather the hash features, 10->16->1 MLP forward +
//                              backward in fp32, run-merged table scatter, weight gradients as bf16 mma
//                              over the 32 samples of a chunk (TC mode) or lane-owned fp32 sums (fp32 mode)
//   tnf_backward_field_kernel_fp32 + tnf_wgrad_kernel_fp32
//                              the exact-arithmetic mode of ThermalNerfactoTField's backward
//                              (thermal_field.py:108-201): lane per sample, fp32 FFMA, (X, dY) rows staged in
//                              global memory for a split-K fp32 weight-gradient pass
//   tensor-core mode           tnf_backward_tc.cu: one kernel, weight gradients accumulated in tensor memory
//                              (tcgen05.mma), nothing per-sample written to HBM
//
// Sample distances are constants here (PDFSampler detaches its bins); with TnfModelGrad.ray_origins /
// ray_directions given, the kernels also return dL/d origins and dL/d directions (camera optimiser).
#include "tnf_backward.cuh"
#include "tnf_host.h"

namespace tnf {

// ------------------------------------------------------------------------------------
// staging layout of the field backward (element type T = float | __nv_bfloat16)
// ------------------------------------------------------------------------------------
struct BwdLayout {
  // per sample rows [Ns, width]
  unsigned char *XF, *XH, *XG, *XA1, *XA2, *XB1, *XB2;   // layer inputs
  unsigned char *dH, *dG, *dGeo, *dA2, *dZ, *dB2, *dT;   // pre-activation gradients
  // per ray rows
  unsigned char *XRay, *dRay;                            // [R,48] sh|appearance, [R,64] sum_s dA1pre
};
constexpr int kWXF = 32, kWXH = 64, kWXG = 16, kWX = 64, kWdGeo = 128, kWdZ = 8, kWdT = 8, kWXRay = 48, kWdRay = 64;

__host__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Carves the staging tensors out of `base` (or only sizes them when L == nullptr).
__host__ inline size_t make_layout(BwdLayout* L, unsigned char* base, long long Ns, long long R, size_t es,
                                   bool own_xf) {
  size_t off = 0;
  auto take = [&](unsigned char* BwdLayout::*field, long long rows, int width) {
    if (L) L->*field = base + off;
    off += align256((size_t)rows * width * es);
  };
  if (own_xf) take(&BwdLayout::XF, Ns, kWXF);
  take(&BwdLayout::XH, Ns, kWXH);
  take(&BwdLayout::XG, Ns, kWXG);
  take(&BwdLayout::XA1, Ns, kWX);
  take(&BwdLayout::XA2, Ns, kWX);
  take(&BwdLayout::XB1, Ns, kWX);
  take(&BwdLayout::XB2, Ns, kWX);
  take(&BwdLayout::dH, Ns, kWX);
  take(&BwdLayout::dG, Ns, kWXG);
  take(&BwdLayout::dGeo, Ns, kWdGeo);
  take(&BwdLayout::dA2, Ns, kWX);
  take(&BwdLayout::dZ, Ns, kWdZ);
  take(&BwdLayout::dB2, Ns, kWX);
  take(&BwdLayout::dT, Ns, kWdT);
  take(&BwdLayout::XRay, R, kWXRay);
  take(&BwdLayout::dRay, R, kWdRay);
  return off;
}

// weight-gradient GEMM problems: dW[n][k] += sum_rows dY[row][n0+n] * X[row][k];  db[n] += sum_rows dY
struct WgradProblem {
  const void* dY; int ldY, n0, N, n_valid;   // N: loaded columns (multiple of 8), n_valid <= N are written
  const void* X;  int ldX, K, k_skip;        // K: loaded columns (multiple of 8); output col = k - k_skip >= 0
  long long rows;
  float* W; int ldW, wcol0;
  float* bias;                               // may be null
};
constexpr int kMaxWgradProblems = 12;
struct WgradArgs {
  WgradProblem p[kMaxWgradProblems];
  int cta_start[kMaxWgradProblems + 1];  // CTAs [cta_start[i], cta_start[i+1]) work on problem i
  int n;
};
// which problem this CTA belongs to, its rank inside the problem's CTA range and that range's size
__device__ __forceinline__ int wgrad_problem_of(const WgradArgs& a, int cta, int& local, int& count) {
  int i = 0;
  while (i + 1 < a.n && cta >= a.cta_start[i + 1]) ++i;
  local = cta - a.cta_start[i];
  count = a.cta_start[i + 1] - a.cta_start[i];
  return i;
}
constexpr int kWgradRows = 64;

// ------------------------------------------------------------------------------------
// proposal levels
//
// Work unit = (level, ray), handed out by a global counter with the 256-sample level first
// (longest-processing-time order), so the two levels of 4096 rays balance over the 2368
// resident warps instead of quantising into whole rays per warp.  Weight gradients of the
// 10->16->1 MLP:
//   TC mode    dW0^T|db0 = dH^T [X | 1] and dW1 = d_o^T relu(H) as bf16 mma.m16n8k16 over the 32
//              samples of a chunk (rows staged in shared memory, fragments via ldmatrix.trans),
//              accumulated in C fragments across all units a warp processes
//   fp32 mode  lane-owned (a,b) column pairs summed over the staged fp32 rows (exact fp32)
// and flushed warp -> CTA (shared atomics) -> global (one atomic per weight per CTA).
// ------------------------------------------------------------------------------------
constexpr int kPropRow = 51;    // fp32 row: dh(16) | do | feat(16) | 1 | relu(h)(16)  (+ pad to an odd stride)
constexpr int kPropSlots = 10;  // ceil((16*16 + 33) / 32)
constexpr int kPropLd = 24;     // bf16 row stride of the staged tiles: 48 B, conflict-free for ldmatrix
constexpr int kPropWacc = 16 * 16 + 33;

struct alignas(16) PropStageTC {
  __nv_bfloat16 dh[32 * kPropLd];  // [sample][j]           dL/dh_pre
  __nv_bfloat16 x[32 * kPropLd];   // [sample][feat.. | 1 at column 2L | 0..]
  __nv_bfloat16 rh[32 * kPropLd];  // [sample][j]           relu(h)
  __nv_bfloat16 dout[32 * 8];      // [sample][d_o, 0 x 7]
};
struct alignas(16) PropStage32 {
  float rows[32 * kPropRow];
};
union PropStage {
  PropStageTC tc;
  PropStage32 f;
};

struct PropBwdScratch {
  float bins[kBuf];
  float gw[kBuf];
  float P[kBuf];  // saved weights, then suffix sums of gw*w
  PropStage st;
};

struct PropBwdSmem {
  PropW prop[TNF_NUM_PROP];
  float wacc[TNF_NUM_PROP][kPropWacc + 3];  // [16*K2 l0.weight | 16 l0.bias | 16 l1.weight | l1.bias]
  PropBwdScratch ws[kWarpsPerCta];
};

// per-warp weight-gradient accumulators of the level being processed
template <bool TC, int NLC>
struct PropAcc;
template <int NLC>
struct PropAcc<true, NLC> {
  static constexpr int NT1 = (2 * NLC + 1 + 7) / 8;  // n-tiles of [X | 1]
  float c1[NT1][4];  // dH^T [X | 1]:  row j, column k (k == 2L: bias)
  float c2[2][4];    // d_o^T relu(H): row 0, column j
  float dosum;       // sum of d_o (l1.bias)
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NT1; ++i) c1[i][0] = c1[i][1] = c1[i][2] = c1[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) c2[i][0] = c2[i][1] = c2[i][2] = c2[i][3] = 0.f;
    dosum = 0.f;
  }
  // warp -> CTA accumulators (layout of PropBwdSmem::wacc)
  __device__ __forceinline__ void flush(float* wacc, const int K2, const int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT1; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = g + (e >> 1) * 8, k = nt * 8 + 2 * q + (e & 1);
        const float v = c1[nt][e];
        if (v != 0.f) {
          if (k < K2) atomicAdd(wacc + j * K2 + k, v);
          else if (k == K2) atomicAdd(wacc + 16 * K2 + j, v);
        }
      }
    if (g == 0) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (c2[nt][e] != 0.f) atomicAdd(wacc + 16 * K2 + 16 + nt * 8 + 2 * q + e, c2[nt][e]);
    }
    const float s = warp_sum(dosum);
    if (lane == 0 && s != 0.f) atomicAdd(wacc + 16 * K2 + 32, s);
  }
};
template <int NLC>
struct PropAcc<false, NLC> {
  float a[kPropSlots];
  int ia[kPropSlots], ib[kPropSlots];  // staged-row columns multiplied by slot r of this lane
  int nslots;
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int r = 0; r < kPropSlots; ++r) a[r] = 0.f;
  }
  __device__ __forceinline__ void make_slots(const int K2, const int lane) {
    const int total = 16 * K2 + 33;
    nslots = (total + 31) / 32;
#pragma unroll
    for (int r = 0; r < kPropSlots; ++r) {
      const int t = lane + 32 * r;
      int a_ = 50, b_ = 50;  // both point at the zeroed pad column -> contributes 0
      if (t < 16 * K2) { a_ = t / K2; b_ = 17 + t % K2; }
      else if (t < 16 * K2 + 16) { a_ = t - 16 * K2; b_ = 33; }
      else if (t < 16 * K2 + 32) { a_ = 16; b_ = 34 + (t - 16 * K2 - 16); }
      else if (t == 16 * K2 + 32) { a_ = 16; b_ = 33; }
      ia[r] = a_;
      ib[r] = b_;
    }
  }
  __device__ __forceinline__ void flush(float* wacc, const int K2, const int lane) {
#pragma unroll
    for (int r = 0; r < kPropSlots; ++r) {
      const int t = lane + 32 * r;
      if (t < 16 * K2 + 33 && a[r] != 0.f) atomicAdd(wacc + t, a[r]);
    }
  }
};

// One (level, ray) unit.  NLC = compile-time bound on the number of hash levels (5 for the nerfacto
// proposal nets, TNF_MAX_PROP_LEVELS otherwise).
template <bool TC, int NLC>
__device__ __forceinline__ void prop_backward_unit(const TnfModel& m, const int lvl, const PropW& W,
                                                   PropBwdScratch& ws, const RayCtx& rc, const int S, const int lane,
                                                   const float* __restrict__ sdist, const float* __restrict__ wsaved,
                                                   const float* __restrict__ gw, float2* __restrict__ gtab,
                                                   PropAcc<TC, NLC>& acc, const bool pose, float (&pg)[6]) {
  const TnfDensityNet& net = m.prop[lvl];
  const int L = net.grid.num_levels;
  const int log2 = net.grid.log2_size;
  const uint32_t mask = (1u << log2) - 1u;
  const float2* __restrict__ tab = reinterpret_cast<const float2*>(net.grid.table);
  for (int i = lane; i <= S; i += 32) ws.bins[i] = sdist[i];
  for (int i = lane; i < S; i += 32) {
    ws.gw[i] = gw[i];
    ws.P[i] = wsaved[i];
  }
  __syncwarp();
  // P_i = sum_{j>i} gw_j * w_j  (chunks visited from the far end)
  {
    float carry = 0.f;
    for (int base = ((S - 1) / 32) * 32; base >= 0; base -= 32) {
      const int i = base + lane;
      const float v = i < S ? ws.gw[i] * ws.P[i] : 0.f;
      float tot;
      const float ex = warp_suffix_excl(v, lane, tot);
      __syncwarp();
      if (i < S) ws.P[i] = ex + carry;
      carry += tot;
    }
    __syncwarp();
  }
  float carry = 0.f;  // sum of delta*sigma of previous chunks
  for (int base = 0; base < S; base += 32) {
    const int i = base + lane;
    const bool active = i < S;
    const int ii = active ? i : S - 1;
    float mid, delta;
    sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
    float px, py, pz;
    const float sel = normalise_position(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), px, py, pz);
    float feat[2 * NLC];
#pragma unroll
    for (int l = 0; l < NLC; ++l) {
      float2 f = make_float2(0.f, 0.f);
      if (l < L) f = hash_level(tab + ((size_t)l << log2), px, py, pz, net.grid.scalings[l], mask);
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
    for (int k = 0; k < 2 * NLC; ++k) {
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
    const float density = expf(o) * sel;
    const float ds = active ? delta * density : 0.f;
    const float incl = warp_incl_scan(ds, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.f;
    const float T = expf(-(carry + excl));
    const float w = (1.f - expf(-ds)) * T;
    carry += __shfl_sync(kFull, incl, 31);
    // d loss / d (delta*sigma), then through density = trunc_exp(o) * selector
    float d_o = 0.f;
    if (active) {
      const float dds = ws.gw[i] * (T - w) - ws.P[i];
      d_o = dds * delta * expf(fminf(fmaxf(o, -15.f), 15.f)) * sel;
    }
    float dh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dh[j] = h[j] > 0.f ? d_o * W.w1[j] : 0.f;

    // ---- stage this chunk's rows for the weight gradients
    if constexpr (TC) {
      PropStageTC& st = ws.st.tc;
      uint4* dhr = reinterpret_cast<uint4*>(st.dh + lane * kPropLd);
      dhr[0] = make_uint4(pack_bf162(dh[0], dh[1]), pack_bf162(dh[2], dh[3]), pack_bf162(dh[4], dh[5]),
                          pack_bf162(dh[6], dh[7]));
      dhr[1] = make_uint4(pack_bf162(dh[8], dh[9]), pack_bf162(dh[10], dh[11]), pack_bf162(dh[12], dh[13]),
                          pack_bf162(dh[14], dh[15]));
      uint4* rhr = reinterpret_cast<uint4*>(st.rh + lane * kPropLd);
      rhr[0] = make_uint4(pack_bf162(fmaxf(h[0], 0.f), fmaxf(h[1], 0.f)), pack_bf162(fmaxf(h[2], 0.f), fmaxf(h[3], 0.f)),
                          pack_bf162(fmaxf(h[4], 0.f), fmaxf(h[5], 0.f)), pack_bf162(fmaxf(h[6], 0.f), fmaxf(h[7], 0.f)));
      rhr[1] = make_uint4(pack_bf162(fmaxf(h[8], 0.f), fmaxf(h[9], 0.f)), pack_bf162(fmaxf(h[10], 0.f), fmaxf(h[11], 0.f)),
                          pack_bf162(fmaxf(h[12], 0.f), fmaxf(h[13], 0.f)), pack_bf162(fmaxf(h[14], 0.f), fmaxf(h[15], 0.f)));
      constexpr int NT1 = PropAcc<true, NLC>::NT1;
      uint32_t xr[NT1 * 4];
#pragma unroll
      for (int c = 0; c < NT1 * 8; c += 2) {
        float v0 = c < 2 * NLC ? feat[c < 2 * NLC ? c : 0] : 0.f;
        float v1 = c + 1 < 2 * NLC ? feat[c + 1 < 2 * NLC ? c + 1 : 0] : 0.f;
        if (c == 2 * L) v0 = 1.f;
        if (c + 1 == 2 * L) v1 = 1.f;
        xr[c >> 1] = pack_bf162(v0, v1);
      }
      uint4* xrow = reinterpret_cast<uint4*>(st.x + lane * kPropLd);
#pragma unroll
      for (int c = 0; c < NT1; ++c) xrow[c] = make_uint4(xr[4 * c], xr[4 * c + 1], xr[4 * c + 2], xr[4 * c + 3]);
      *reinterpret_cast<uint4*>(st.dout + lane * 8) = make_uint4(pack_bf162(d_o, 0.f), 0u, 0u, 0u);
    } else {
      float* row = ws.st.f.rows + lane * kPropRow;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        row[j] = dh[j];
        row[34 + j] = fmaxf(h[j], 0.f);
      }
      row[16] = d_o;
      row[33] = 1.f;
      row[50] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) row[17 + k] = k < 2 * NLC ? feat[k < 2 * NLC ? k : 0] : 0.f;
    }

    // ---- table scatter: dfeat = dh . W0 (whole chunk skipped when no sample of it carries a gradient)
    if (__any_sync(kFull, d_o != 0.f)) {
      float dpx = 0.f, dpy = 0.f, dpz = 0.f;
#pragma unroll
      for (int l = 0; l < NLC; ++l) {
        if (l < L) {
          float gx = 0.f, gy = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            gx = fmaf(dh[j], W.w0t[(2 * l) * 16 + j], gx);
            gy = fmaf(dh[j], W.w0t[(2 * l + 1) * 16 + j], gy);
          }
          scatter_level_runs<1>(gtab + ((size_t)l << log2), px, py, pz, net.grid.scalings[l], mask, gx, gy, lane);
          if (pose && d_o != 0.f)
            hash_level_pos_grad(tab + ((size_t)l << log2), px, py, pz, net.grid.scalings[l], mask, gx, gy, dpx, dpy, dpz);
        }
      }
      if (pose && d_o != 0.f) {
        position_grad_to_world(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), sel, dpx, dpy, dpz);
        pg[0] += dpx; pg[1] += dpy; pg[2] += dpz;
        pg[3] = fmaf(mid, dpx, pg[3]); pg[4] = fmaf(mid, dpy, pg[4]); pg[5] = fmaf(mid, dpz, pg[5]);
      }
    }
    __syncwarp();

    // ---- weight gradients of this chunk
    if constexpr (TC) {
      PropStageTC& st = ws.st.tc;
      constexpr int NT1 = PropAcc<true, NLC>::NT1;
      const int mi = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, &st.dh[(ks * 16 + (mi >> 1) * 8 + r) * kPropLd + (mi & 1) * 8]);
        {
          uint32_t b[4];
          ldmatrix_x4_trans(b, &st.x[(ks * 16 + (mi & 1) * 8 + r) * kPropLd + (mi >> 1) * 8]);
          mma_16816_bf16(acc.c1[0], a, make_uint2(b[0], b[1]));
          mma_16816_bf16(acc.c1[1], a, make_uint2(b[2], b[3]));
        }
        if constexpr (NT1 > 2) {
          uint32_t b[2];
          ldmatrix_x2_trans(b, &st.x[(ks * 16 + (mi & 1) * 8 + r) * kPropLd + 16]);
          mma_16816_bf16(acc.c1[NT1 - 1], a, make_uint2(b[0], b[1]));
        }
        uint32_t a2[4], t2[2];
        ldmatrix_x2_trans(t2, &st.dout[(ks * 16 + (mi & 1) * 8 + r) * 8]);
        a2[0] = t2[0]; a2[1] = 0u; a2[2] = t2[1]; a2[3] = 0u;
        uint32_t b2[4];
        ldmatrix_x4_trans(b2, &st.rh[(ks * 16 + (mi & 1) * 8 + r) * kPropLd + (mi >> 1) * 8]);
        mma_16816_bf16(acc.c2[0], a2, make_uint2(b2[0], b2[1]));
        mma_16816_bf16(acc.c2[1], a2, make_uint2(b2[2], b2[3]));
      }
      acc.dosum += d_o;
    } else {
      for (int s = 0; s < 32; ++s) {
        const float* r_ = ws.st.f.rows + s * kPropRow;
#pragma unroll
        for (int r = 0; r < kPropSlots; ++r)
          if (r < acc.nslots) acc.a[r] = fmaf(r_[acc.ia[r]], r_[acc.ib[r]], acc.a[r]);
      }
    }
    __syncwarp();
  }
}

template <bool TC, int NLC>
__global__ void __launch_bounds__(kThreads, 2)
    tnf_backward_prop_kernel(const __grid_constant__ TnfModel m, const __grid_constant__ TnfRays rays,
                             const __grid_constant__ TnfSaved sv, const __grid_constant__ TnfOutputGrads go,
                             const __grid_constant__ TnfModelGrad gr, unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PropBwdSmem& S = *reinterpret_cast<PropBwdSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_prop(S.prop[0], m.prop[0], tid);
  stage_prop(S.prop[1], m.prop[1], tid);
  for (int i = tid; i < TNF_NUM_PROP * (kPropWacc + 3); i += kThreads) (&S.wacc[0][0])[i] = 0.f;
  __syncthreads();
  PropBwdScratch& ws = S.ws[warp];
  const bool on[TNF_NUM_PROP] = {gr.prop[0].table != nullptr && go.weights[0] != nullptr,
                                 gr.prop[1].table != nullptr && go.weights[1] != nullptr};
  const unsigned long long R = (unsigned long long)rays.num_rays;
  const unsigned long long n0 = on[0] ? R : 0ull, ntasks = n0 + (on[1] ? R : 0ull);
  PropAcc<TC, NLC> acc;
  acc.zero();
  int cur = -1;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(counter, 1ull);
    t = __shfl_sync(kFull, t, 0);
    if (t >= ntasks) break;
    const int lvl = t < n0 ? 0 : 1;
    const long long ray = (long long)(t < n0 ? t : t - n0);
    if (lvl != cur) {
      if (cur >= 0) acc.flush(S.wacc[cur], 2 * m.prop[cur].grid.num_levels, lane);
      acc.zero();
      if constexpr (!TC) acc.make_slots(2 * m.prop[lvl].grid.num_levels, lane);
      cur = lvl;
    }
    RayCtx rc;
    rc.ox = __ldg(rays.origins + ray * 3 + 0);
    rc.oy = __ldg(rays.origins + ray * 3 + 1);
    rc.oz = __ldg(rays.origins + ray * 3 + 2);
    rc.dx = __ldg(rays.directions + ray * 3 + 0);
    rc.dy = __ldg(rays.directions + ray * 3 + 1);
    rc.dz = __ldg(rays.directions + ray * 3 + 2);
    rc.s_near = spacing_fn(rays.nears ? __ldg(rays.nears + ray) : m.near_plane);
    rc.s_far = spacing_fn(rays.fars ? __ldg(rays.fars + ray) : m.far_plane);
    const int S_l = m.num_samples[lvl];
    float pg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    prop_backward_unit<TC, NLC>(m, lvl, S.prop[lvl], ws, rc, S_l, lane, sv.sdist[lvl] + ray * (S_l + 1),
                                sv.weights[lvl] + ray * S_l, go.weights[lvl] + ray * S_l,
                                reinterpret_cast<float2*>(gr.prop[lvl].table), acc, gr.ray_origins != nullptr, pg);
    if (gr.ray_origins) flush_ray_grad(gr, ray, pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], lane);
  }
  if (cur >= 0) acc.flush(S.wacc[cur], 2 * m.prop[cur].grid.num_levels, lane);
  __syncthreads();
  // CTA -> global
#pragma unroll
  for (int lvl = 0; lvl < TNF_NUM_PROP; ++lvl) {
    if (!on[lvl]) continue;
    const int K2 = 2 * m.prop[lvl].grid.num_levels;
    const TnfDensityNetGrad& g = gr.prop[lvl];
    for (int t = tid; t < 16 * K2 + 33; t += kThreads) {
      const float v = S.wacc[lvl][t];
      if (v == 0.f) continue;
      if (t < 16 * K2) atomicAdd(g.l0.weight + t, v);  // [16, K2] row-major: j*K2 + k == t
      else if (t < 16 * K2 + 16) atomicAdd(g.l0.bias + (t - 16 * K2), v);
      else if (t < 16 * K2 + 32) atomicAdd(g.l1.weight + (t - 16 * K2 - 16), v);
      else atomicAdd(g.l1.bias, v);
    }
  }
}

// per-ray epilogue: stage [sh | appearance] and sum_s dA1pre, appearance-embedding gradient
template <typename T>
__device__ __forceinline__ void ray_epilogue(const TnfModel& m, const TnfRays& rays, const long long ray,
                                             const int lane, const float (&sh)[16], const float app_lane,
                                             const float racc0, const float racc1, const BwdLayout& L,
                                             float* __restrict__ gapp) {
  T* xr = reinterpret_cast<T*>(L.XRay) + ray * kWXRay;
  T* dr = reinterpret_cast<T*>(L.dRay) + ray * kWdRay;
  if (lane < 16) xr[lane] = T(sh[lane]);
  xr[16 + lane] = T(app_lane);
  dr[lane] = T(racc0);
  dr[lane + 32] = T(racc1);
  if (m.appearance_mode == TNF_APPEARANCE_LOOKUP && gapp) {
    const float* W = m.field.rgb0.weight;
    float acc = 0.f;
    for (int n = 0; n < 64; ++n) {
      const float v = __shfl_sync(kFull, n < 32 ? racc0 : racc1, n & 31);
      acc = fmaf(v, W[n * 63 + 31 + lane], acc);
    }
    atomicAdd(gapp + rays.camera_indices[ray] * 32 + lane, acc);
  }
}

// ------------------------------------------------------------------------------------
// field level, fp32: lane per sample (mirror of the fp32 forward)
// ------------------------------------------------------------------------------------
struct FieldBwdSmem32 {
  FieldW32 fw;
  FieldBwdScratch ws[kWarpsPerCta];
  float act[kWarpsPerCta][64 * 32];
  float geo[kWarpsPerCta][16 * 32];
  float dgc[kWarpsPerCta][16 * 32];
};

// out[k] = sum_n dy[n] * wt[k*N + n]   (wt is the k-major image of torch weight[n][k]);
// MASK: multiply by (col[k] > 0) where col currently holds the layer input; result replaces col[k].
template <int N, int K, bool MASK, bool ACCUM>
__device__ __forceinline__ void dense_row_t(const float* __restrict__ wt, const float (&dy)[N], float* col) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int n = 0; n < N; n += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + k * N + n);
      a0 = fmaf(dy[n], w.x, a0);
      a1 = fmaf(dy[n + 1], w.y, a1);
      a2 = fmaf(dy[n + 2], w.z, a2);
      a3 = fmaf(dy[n + 3], w.w, a3);
    }
    float r = (a0 + a1) + (a2 + a3);
    if (MASK) r = col[k * 32] > 0.f ? r : 0.f;
    if (ACCUM) r += col[k * 32];
    col[k * 32] = r;
  }
}
template <int N>
__device__ __forceinline__ void load_col(const float* col, float (&y)[N]) {
#pragma unroll
  for (int n = 0; n < N; ++n) y[n] = col[n * 32];
}
template <int N>
__device__ __forceinline__ void stage_row(unsigned char* base, long long row, int ld, int col0, const float (&y)[N]) {
  float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + row * ld + col0);
#pragma unroll
  for (int n = 0; n < N; n += 4) p[n >> 2] = make_float4(y[n], y[n + 1], y[n + 2], y[n + 3]);
}

__global__ void __launch_bounds__(kThreads, 1)
    tnf_backward_field_kernel_fp32(const __grid_constant__ TnfModel m, const __grid_constant__ TnfRays rays,
                                   const __grid_constant__ TnfSaved sv, const __grid_constant__ TnfOutputGrads go,
                                   const __grid_constant__ TnfModelGrad gr, const __grid_constant__ BwdLayout L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FieldBwdSmem32& S = *reinterpret_cast<FieldBwdSmem32*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_field(S.fw, m.field, tid, kThreads, m.head_mode == TNF_HEAD_CONCAT ? 4 : 3);
  __syncthreads();
  const FieldW32& W = S.fw;
  FieldBwdScratch& ws = S.ws[warp];
  float* a = &S.act[warp][lane];
  float* geo = &S.geo[warp][lane];
  float* dgc = &S.dgc[warp][lane];
  const int S2 = m.num_samples[TNF_NUM_PROP];
  const TnfHashGrid& grid = m.field.grid;
  const uint32_t mask = (1u << grid.log2_size) - 1u;
  float2* __restrict__ gtab = reinterpret_cast<float2*>(gr.field.table);
  const float* __restrict__ F = static_cast<const float*>(sv.field_features);
  const long long R = rays.num_rays;

  for (long long ray = (long long)blockIdx.x * kWarpsPerCta + warp; ray < R;
       ray += (long long)gridDim.x * kWarpsPerCta) {
    RayCtx rc;
    rc.ox = __ldg(rays.origins + ray * 3 + 0);
    rc.oy = __ldg(rays.origins + ray * 3 + 1);
    rc.oz = __ldg(rays.origins + ray * 3 + 2);
    rc.dx = __ldg(rays.directions + ray * 3 + 0);
    rc.dy = __ldg(rays.directions + ray * 3 + 1);
    rc.dz = __ldg(rays.directions + ray * 3 + 2);
    rc.s_near = spacing_fn(rays.nears ? __ldg(rays.nears + ray) : m.near_plane);
    rc.s_far = spacing_fn(rays.fars ? __ldg(rays.fars + ray) : m.far_plane);
    for (int i = lane; i <= S2; i += 32) ws.bins[i] = sv.sdist[TNF_NUM_PROP][ray * (S2 + 1) + i];
    float sh[16], app_lane;
    ray_bias_and_inputs(m, rays, ray, rc, lane, ws.rayb, sh, app_lane);
    __syncwarp();
    composite_backward(m, ws, rc, S2, lane, sv.field_samples + ray * S2 * 5,
                       go.weights[TNF_NUM_PROP] ? go.weights[TNF_NUM_PROP] + ray * S2 : nullptr,
                       go.rgb ? go.rgb[ray * 3 + 0] : 0.f, go.rgb ? go.rgb[ray * 3 + 1] : 0.f,
                       go.rgb ? go.rgb[ray * 3 + 2] : 0.f, go.thermal ? go.thermal[ray] : 0.f,
                       go.accumulation ? go.accumulation[ray] : 0.f);
    float racc0 = 0.f, racc1 = 0.f;
    float pg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool pose = gr.ray_origins != nullptr;
    for (int base = 0; base < S2; base += 32) {
      const int i = base + lane;
      const bool active = i < S2;
      const int ii = active ? i : S2 - 1;
      const long long row = ray * S2 + ii;
      float mid, delta;
      sample_geometry(rc, ws.bins[ii], ws.bins[ii + 1], mid, delta);
      float px, py, pz;
      const float sel = normalise_position(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), px, py, pz);
      // upstream gradients of this sample (zero for padding lanes: every product below vanishes)
      const float dsig = active ? ws.dsig[ii] : 0.f;
      // ws.dtau: thermal head -> dL/d thermal_i; RGBT head (concat_nerf) -> dL/d(pre-sigmoid channel 3)
      const bool concat = m.head_mode == TNF_HEAD_CONCAT;
      const float dt = active ? ws.dtau[ii] : 0.f;
      const float dz[4] = {active ? ws.dzr[ii] : 0.f, active ? ws.dzg[ii] : 0.f, active ? ws.dzb[ii] : 0.f,
                           concat ? dt : 0.f};
      const float dtau = concat ? 0.f : dt;
      // ---- forward recompute: F -> H -> G
      {
        const float4* fr = reinterpret_cast<const float4*>(F + row * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 v = fr[k];
          a[(4 * k) * 32] = v.x; a[(4 * k + 1) * 32] = v.y; a[(4 * k + 2) * 32] = v.z; a[(4 * k + 3) * 32] = v.w;
        }
      }
      float h0;
      {
        float y[64];
        dense_col<32, 64, ACT_RELU>(W.base0t, W.base0b, a, y);
        store_col(a, y);
        if (active) stage_row(L.XH, row, kWXH, 0, y);
      }
      {
        float y[16];
        dense_col<64, 16, ACT_NONE>(W.base1t, W.base1b, a, y);
        h0 = y[0];
        if (active) stage_row(L.XG, row, kWXG, 0, y);
        y[0] = 0.f;
        store_col(geo, y);
      }
      // ---- thermal head: forward then backward down to dB1pre and its share of dG
      {
        float y[64];
        dense_col<16, 64, ACT_RELU>(W.th0t, W.th0b, geo, y);
        store_col(a, y);
        if (active) stage_row(L.XB1, row, kWX, 0, y);
        dense_col<64, 64, ACT_SIGMOID>(W.th1t, W.th1b, a, y);
        if (active) stage_row(L.XB2, row, kWX, 0, y);
#pragma unroll
        for (int n = 0; n < 64; ++n) y[n] = dtau * W.th2[n] * y[n] * (1.f - y[n]);
        if (active) stage_row(L.dB2, row, kWX, 0, y);
        dense_row_t<64, 64, true, false>(W.th1t, y, a);  // a: B1 -> dB1pre
        load_col(a, y);
        if (active) stage_row(L.dGeo, row, kWdGeo, 64, y);
        if (m.detach_thermal_geo) {
#pragma unroll
          for (int k = 0; k < 16; ++k) dgc[k * 32] = 0.f;
        } else {
          dense_row_t<64, 16, false, false>(W.th0t, y, dgc);
        }
        float t8[8] = {dtau, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (active) stage_row(L.dT, row, kWdT, 0, t8);
      }
      // ---- colour head
      {
        float y[64];
        dense_col<16, 64, ACT_RELU>(W.rgb0geo_t, ws.rayb, geo, y);
        store_col(a, y);
        if (active) stage_row(L.XA1, row, kWX, 0, y);
        dense_col<64, 64, ACT_RELU>(W.rgb1t, W.rgb1b, a, y);
        if (active) stage_row(L.XA2, row, kWX, 0, y);
#pragma unroll
        for (int n = 0; n < 64; ++n) {
          const float4 w2 = *reinterpret_cast<const float4*>(W.rgb2t + n * 4);
          y[n] = y[n] > 0.f ? dz[0] * w2.x + dz[1] * w2.y + dz[2] * w2.z + dz[3] * w2.w : 0.f;
        }
        if (active) stage_row(L.dA2, row, kWX, 0, y);
        float z8[8] = {dz[0], dz[1], dz[2], dz[3], 0.f, 0.f, 0.f, 0.f};
        if (active) stage_row(L.dZ, row, kWdZ, 0, z8);
        dense_row_t<64, 64, true, false>(W.rgb1t, y, a);  // a: A1 -> dA1pre
        load_col(a, y);
        if (active) stage_row(L.dGeo, row, kWdGeo, 0, y);
#pragma unroll
        for (int n = 0; n < 64; ++n) {
          const float s = warp_sum(y[n]);
          if (lane == (n & 31)) { if (n < 32) racc0 += s; else racc1 += s; }
        }
        dense_row_t<64, 16, false, true>(W.rgb0geo_t, y, dgc);
      }
      // ---- trunk: dG -> dH -> dF
      {
        float y[16];
        load_col(dgc, y);
        y[0] = dsig * expf(fminf(fmaxf(h0, -15.f), 15.f)) * sel;  // trunc_exp backward, times selector
        if (active) stage_row(L.dG, row, kWXG, 0, y);
        {
          const float4* hr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(L.XH) + row * kWXH);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float4 v = hr[k];
            a[(4 * k) * 32] = v.x; a[(4 * k + 1) * 32] = v.y; a[(4 * k + 2) * 32] = v.z; a[(4 * k + 3) * 32] = v.w;
          }
        }
        dense_row_t<16, 64, true, false>(W.base1t, y, a);  // a: H -> dHpre
      }
      {
        float y[64];
        load_col(a, y);
        if (active) stage_row(L.dH, row, kWX, 0, y);
        dense_row_t<64, 32, false, false>(W.base0t, y, a);  // a[0..31]: dF
      }
      if (active) {
        float dpx = 0.f, dpy = 0.f, dpz = 0.f;
#pragma unroll 4
        for (int l = 0; l < TNF_MAX_LEVELS; ++l) {
          scatter_level(gtab + ((size_t)l << grid.log2_size), px, py, pz, grid.scalings[l], mask, a[(2 * l) * 32],
                        a[(2 * l + 1) * 32]);
          if (pose)
            hash_level_pos_grad(reinterpret_cast<const float2*>(grid.table) + ((size_t)l << grid.log2_size), px, py, pz,
                                grid.scalings[l], mask, a[(2 * l) * 32], a[(2 * l + 1) * 32], dpx, dpy, dpz);
        }
        if (pose) {
          position_grad_to_world(m, ray_x(rc, mid), ray_y(rc, mid), ray_z(rc, mid), sel, dpx, dpy, dpz);
          pg[0] += dpx; pg[1] += dpy; pg[2] += dpz;
          pg[3] = fmaf(mid, dpx, pg[3]); pg[4] = fmaf(mid, dpy, pg[4]); pg[5] = fmaf(mid, dpz, pg[5]);
        }
      }
    }
    if (pose) flush_ray_grad(gr, ray, pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], lane);
    ray_epilogue<float>(m, rays, ray, lane, sh, app_lane, racc0, racc1, L, gr.field.appearance);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// weight-gradient GEMMs: dW[n][k] += sum_rows dY[row][n0+n] * X[row][k];  db[n] += sum_rows dY
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tnf_wgrad_kernel_fp32(const __grid_constant__ WgradArgs args) {
  int cta_local, cta_count;
  const WgradProblem& P = args.p[wgrad_problem_of(args, blockIdx.x, cta_local, cta_count)];
  __shared__ __align__(16) float sdY[kWgradRows][68];
  __shared__ __align__(16) float sX[kWgradRows][68];
  const int tid = threadIdx.x, tn = tid >> 4, tk = tid & 15;
  const float* dY = static_cast<const float*>(P.dY);
  const float* X = static_cast<const float*>(P.X);
  float acc[4][4] = {};
  float bacc[4] = {};
  const long long tiles = (P.rows + kWgradRows - 1) / kWgradRows;
  for (long long t = cta_local; t < tiles; t += cta_count) {
    const long long row0 = t * kWgradRows;
    for (int idx = tid; idx < kWgradRows * 16; idx += 256) {
      const int r = idx >> 4, c4 = (idx & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), x = v;
      if (row0 + r < P.rows) {
        if (c4 < P.N) v = *reinterpret_cast<const float4*>(dY + (row0 + r) * P.ldY + P.n0 + c4);
        if (c4 < P.K) x = *reinterpret_cast<const float4*>(X + (row0 + r) * P.ldX + c4);
      }
      *reinterpret_cast<float4*>(&sdY[r][c4]) = v;
      *reinterpret_cast<float4*>(&sX[r][c4]) = x;
    }
    __syncthreads();
    if (4 * tn < P.N && 4 * tk < P.K) {
#pragma unroll 4
      for (int r = 0; r < kWgradRows; ++r) {
        const float4 d = *reinterpret_cast<const float4*>(&sdY[r][4 * tn]);
        const float4 x = *reinterpret_cast<const float4*>(&sX[r][4 * tk]);
        const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
          if (tk == 0) bacc[i] += dv[i];
        }
      }
    }
    __syncthreads();
  }
  if (4 * tn < P.N && 4 * tk < P.K) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = 4 * tn + i;
      if (n >= P.n_valid) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * tk + j - P.k_skip;
        if (k >= 0 && 4 * tk + j < P.K && acc[i][j] != 0.f) atomicAdd(P.W + n * P.ldW + P.wcol0 + k, acc[i][j]);
      }
      if (tk == 0 && P.bias && bacc[i] != 0.f) atomicAdd(P.bias + n, bacc[i]);
    }
  }
}

}  // namespace tnf

// ====================================================================================
// C ABI
// ====================================================================================
namespace {
using tnf::fail;

int check_grads(const TnfModel& m, const TnfModelGrad* g) {
  if (!g) return fail(TNF_ERR_INVALID_ARGUMENT, "grads is null");
  for (int k = 0; k < TNF_NUM_PROP; ++k) {
    const TnfDensityNetGrad& p = g->prop[k];
    if (p.table && (!p.l0.weight || !p.l0.bias || !p.l1.weight || !p.l1.bias))
      return fail(TNF_ERR_INVALID_ARGUMENT, "grads.prop[%d]: table given but a linear gradient is null", k);
  }
  const TnfFieldGrad& f = g->field;
  if (!f.table) return fail(TNF_ERR_INVALID_ARGUMENT, "grads.field.table is null");
  const TnfLinearGrad* ls[] = {&f.base0, &f.base1, &f.rgb0, &f.rgb1, &f.rgb2, &f.th0, &f.th1, &f.th2};
  for (const TnfLinearGrad* l : ls)
    if (!l->weight || !l->bias) return fail(TNF_ERR_INVALID_ARGUMENT, "grads.field: a linear gradient is null");
  if (m.appearance_mode == TNF_APPEARANCE_LOOKUP && !f.appearance)
    return fail(TNF_ERR_INVALID_ARGUMENT, "grads.field.appearance is required for TNF_APPEARANCE_LOOKUP");
  if ((g->ray_origins == nullptr) != (g->ray_directions == nullptr))
    return fail(TNF_ERR_INVALID_ARGUMENT, "grads.ray_origins and grads.ray_directions must be given together");
  return TNF_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaFuncSetAttribute(%s, smem=%zu): %s", name, bytes,
                                    cudaGetErrorString(e));
  return TNF_OK;
}

void add_problem(tnf::WgradArgs& a, const void* dY, int ldY, int n0, int N, int n_valid, const void* X, int ldX,
                 int K, int k_skip, long long rows, float* W, int ldW, int wcol0, float* bias) {
  tnf::WgradProblem& p = a.p[a.n++];
  p.dY = dY; p.ldY = ldY; p.n0 = n0; p.N = N; p.n_valid = n_valid;
  p.X = X; p.ldX = ldX; p.K = K; p.k_skip = k_skip;
  p.rows = rows; p.W = W; p.ldW = ldW; p.wcol0 = wcol0; p.bias = bias;
}

// CTAs per problem in proportion to the bytes it streams (rows x loaded columns): the 64x64 layers move 1.7x the
// average, and an equal split made the whole launch wait for them.  Returns the grid size.
int assign_wgrad_ctas(tnf::WgradArgs& a, int total_ctas) {
  double cost[tnf::kMaxWgradProblems], sum = 0.0;
  for (int i = 0; i < a.n; ++i) {
    cost[i] = (double)a.p[i].rows * (a.p[i].N + a.p[i].K);
    sum += cost[i];
  }
  int acc = 0;
  for (int i = 0; i < a.n; ++i) {
    a.cta_start[i] = acc;
    const long long tiles = (a.p[i].rows + tnf::kWgradRows - 1) / tnf::kWgradRows;
    long long want = (long long)(cost[i] / sum * total_ctas + 0.5);
    if (want < 1) want = 1;
    if (want > tiles) want = tiles > 0 ? tiles : 1;
    acc += (int)want;
  }
  a.cta_start[a.n] = acc;
  return acc;
}

// the eight field layers (+ the two per-ray blocks of mlp_head.layers.0) as GEMM problems of the fp32 mode:
// plain row-major fp32 staging, dGeo = [Ns,128] (colour | thermal layer-0 gradients).
void field_problems(tnf::WgradArgs& a, const tnf::BwdLayout& L, const void* XF, const TnfFieldGrad& g, long long Ns,
                    long long R, int nout) {
  using namespace tnf;
  a.n = 0;
  auto off = [&](const unsigned char* p, size_t elems) { return static_cast<const void*>(p + elems * 4); };
  add_problem(a, L.dH, kWX, 0, 64, 64, XF, kWXF, 32, 0, Ns, g.base0.weight, 32, 0, g.base0.bias);
  add_problem(a, L.dG, kWXG, 0, 16, 16, L.XH, kWXH, 64, 0, Ns, g.base1.weight, 64, 0, g.base1.bias);
  add_problem(a, L.dGeo, kWdGeo, 0, 64, 64, L.XG, kWXG, 16, 1, Ns, g.rgb0.weight, 63, 16, nullptr);
  add_problem(a, L.dGeo, kWdGeo, 64, 64, 64, L.XG, kWXG, 16, 1, Ns, g.th0.weight, 15, 0, g.th0.bias);
  add_problem(a, L.dA2, kWX, 0, 64, 64, L.XA1, kWX, 64, 0, Ns, g.rgb1.weight, 64, 0, g.rgb1.bias);
  add_problem(a, L.dZ, kWdZ, 0, 8, nout, L.XA2, kWX, 64, 0, Ns, g.rgb2.weight, 64, 0, g.rgb2.bias);
  add_problem(a, L.dB2, kWX, 0, 64, 64, L.XB1, kWX, 64, 0, Ns, g.th1.weight, 64, 0, g.th1.bias);
  add_problem(a, L.dT, kWdT, 0, 8, 1, L.XB2, kWX, 64, 0, Ns, g.th2.weight, 64, 0, g.th2.bias);
  add_problem(a, L.dRay, kWdRay, 0, 64, 64, L.XRay, kWXRay, 16, 0, R, g.rgb0.weight, 63, 0, g.rgb0.bias);
  add_problem(a, L.dRay, kWdRay, 0, 64, 64, off(L.XRay, 16), kWXRay, 32, 0, R, g.rgb0.weight, 63, 31, nullptr);
}
}  // namespace

namespace {
thread_local int g_stage_mask = 7;
}

extern "C" {

int tnf_backward_stage_mask(int mask) {
  const int prev = g_stage_mask;
  g_stage_mask = mask & 7;
  return prev;
}

size_t tnf_backward_workspace_bytes(const TnfModel* model, int64_t num_rays) {
  if (!model || num_rays <= 0) return 256;
  const long long Ns = (long long)num_rays * model->num_samples[TNF_NUM_PROP];
  // tensor-core mode keeps every per-sample intermediate on chip (tnf_backward_tc.cu): only the work counter of
  // the proposal kernel lives here.  fp32 mode stages the (X, dY) rows of the weight-gradient pass.
  if (model->precision == TNF_PRECISION_TC_FP16) return 256;
  return tnf::make_layout(nullptr, nullptr, Ns, num_rays, 4, false) + 256;
}

int tnf_render_backward(const TnfModel* model, const TnfRays* rays, const TnfSaved* saved, const TnfOutputGrads* gout,
                        const TnfModelGrad* grads, void* workspace, size_t workspace_bytes, void* stream_) {
  return tnf_render_backward_staged(model, rays, saved, gout, grads, workspace, workspace_bytes, stream_, nullptr, 0);
}

int tnf_render_backward_staged(const TnfModel* model, const TnfRays* rays, const TnfSaved* saved,
                               const TnfOutputGrads* gout, const TnfModelGrad* grads, void* workspace,
                               size_t workspace_bytes, void* stream_, void* field_grads_done, int32_t reserve_ctas) {
  tnf::g_err[0] = 0;
  if (int e = tnf::check_model(model)) return e;
  if (!rays || !saved || !gout) return fail(TNF_ERR_INVALID_ARGUMENT, "rays/saved/gout is null");
  if (int e = check_grads(*model, grads)) return e;
  if (!model->training) return fail(TNF_ERR_INVALID_ARGUMENT, "backward needs a training-mode forward (training=1)");
  const long long R = rays->num_rays;
  if (R < 0) return fail(TNF_ERR_INVALID_ARGUMENT, "num_rays=%lld", R);
  if (R == 0) return TNF_OK;
  if (rays->from_camera) return fail(TNF_ERR_INVALID_ARGUMENT, "from_camera rays are an eval-mode input");
  if (!rays->origins || !rays->directions) return fail(TNF_ERR_INVALID_ARGUMENT, "origins/directions is null");
  if (model->appearance_mode == TNF_APPEARANCE_LOOKUP && !rays->camera_indices)
    return fail(TNF_ERR_INVALID_ARGUMENT, "camera_indices required for TNF_APPEARANCE_LOOKUP");
  for (int k = 0; k <= TNF_NUM_PROP; ++k) {
    const bool need = k == TNF_NUM_PROP || (grads->prop[k].table && gout->weights[k]);
    if (need && (!saved->sdist[k] || !saved->weights[k]))
      return fail(TNF_ERR_INVALID_ARGUMENT, "saved.sdist[%d]/weights[%d] is null", k, k);
  }
  if (!saved->field_features || !saved->field_samples)
    return fail(TNF_ERR_INVALID_ARGUMENT, "saved.field_features/field_samples is null");
  const size_t need = tnf_backward_workspace_bytes(model, R);
  if (!workspace || workspace_bytes < need)
    return fail(TNF_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu bytes", workspace_bytes, need);
  if (!tnf::aligned16(workspace)) return fail(TNF_ERR_INVALID_ARGUMENT, "workspace must be 16-byte aligned");

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int sms = tnf::num_sms();
  const long long want = (R + tnf::kWarpsPerCta - 1) / tnf::kWarpsPerCta;
  cudaError_t e;

  // ---- proposal levels: (level, ray) units handed out by a device counter (last 8 bytes of the workspace)
  const bool do_prop = (grads->prop[0].table && gout->weights[0]) || (grads->prop[1].table && gout->weights[1]);
  auto launch_prop = [&]() -> int {
    if (!(do_prop && (g_stage_mask & 1))) return TNF_OK;
    unsigned long long* counter =
        reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + need - 16);
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const size_t smem = sizeof(tnf::PropBwdSmem);
    const bool tc_ = model->precision == TNF_PRECISION_TC_FP16;
    const bool five = model->prop[0].grid.num_levels <= 5 && model->prop[1].grid.num_levels <= 5;
    long long units = 0;
    for (int k = 0; k < TNF_NUM_PROP; ++k)
      if (grads->prop[k].table && gout->weights[k]) units += R;
    const long long wantu = (units + tnf::kWarpsPerCta - 1) / tnf::kWarpsPerCta;
    // two CTAs per SM fill the register file; `reserve_ctas` slots of 256 threads x 64 registers (half a CTA of this
    // kernel each) stay free for a kernel on another stream.  Units come from a device counter: fewer CTAs just
    // take more units each.
    long long cap = (long long)sms * 2 - (reserve_ctas > 0 ? (reserve_ctas + 1) / 2 : 0);
    if (cap < sms) cap = sms;
    const unsigned grid = (unsigned)(wantu < cap ? wantu : cap);
#define TNF_LAUNCH_PROP(TC_, NLC_)                                                                              \
  do {                                                                                                          \
    if (int rc = set_smem(tnf::tnf_backward_prop_kernel<TC_, NLC_>, smem, "backward_prop")) return rc;          \
    tnf::tnf_backward_prop_kernel<TC_, NLC_><<<grid, tnf::kThreads, smem, stream>>>(*model, *rays, *saved,      \
                                                                                   *gout, *grads, counter);     \
  } while (0)
    if (tc_ && five) TNF_LAUNCH_PROP(true, 5);
    else if (tc_) TNF_LAUNCH_PROP(true, TNF_MAX_PROP_LEVELS);
    else if (five) TNF_LAUNCH_PROP(false, 5);
    else TNF_LAUNCH_PROP(false, TNF_MAX_PROP_LEVELS);
#undef TNF_LAUNCH_PROP
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "backward_prop launch: %s", cudaGetErrorString(e));
    return TNF_OK;
  };
  // The two levels write disjoint gradients (the interlevel loss sees the field's weights detached).  Default order:
  // proposal, field.  Staged: field first, the event recorded, then proposal - the caller starts exchanging the field
  // gradients with the other GPUs while the proposal kernel still runs.
  auto field_done = [&]() -> int {
    if (!field_grads_done) return TNF_OK;
    const cudaError_t re = cudaEventRecord(static_cast<cudaEvent_t>(field_grads_done), stream);
    if (re != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(re));
    return launch_prop();
  };
  if (!field_grads_done)
    if (int rc = launch_prop()) return rc;

  // ---- field level
  const long long Ns = R * model->num_samples[TNF_NUM_PROP];
  if (model->precision == TNF_PRECISION_TC_FP16) {
    if (g_stage_mask & 2)
      if (int rc = tnf::launch_backward_field_tc(*model, *rays, *saved, *gout, *grads, stream)) return rc;
    return field_done();
  }
  tnf::BwdLayout L{};
  tnf::make_layout(&L, static_cast<unsigned char*>(workspace), Ns, R, 4, false);
  tnf::WgradArgs wa;
  const size_t smem = sizeof(tnf::FieldBwdSmem32);
  if (int rc = set_smem(tnf::tnf_backward_field_kernel_fp32, smem, "backward_field_fp32")) return rc;
  const long long cap = sms;
  if (g_stage_mask & 2)
    tnf::tnf_backward_field_kernel_fp32<<<(unsigned)(want < cap ? want : cap), tnf::kThreads, smem, stream>>>(
        *model, *rays, *saved, *gout, *grads, L);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "backward_field launch: %s", cudaGetErrorString(e));
  field_problems(wa, L, saved->field_features, grads->field, Ns, R, model->head_mode == TNF_HEAD_CONCAT ? 4 : 3);
  const int grid = assign_wgrad_ctas(wa, sms * 2);
  if (g_stage_mask & 4) tnf::tnf_wgrad_kernel_fp32<<<grid, 256, 0, stream>>>(wa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(e));
  return field_done();
}

}  // extern "C"
