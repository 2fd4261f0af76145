// Fused losses of ThermalNerfModel.get_loss_dict
// (thermo_nerf/thermal_nerf/thermal_nerf_model.py:277-326) and their gradients:
//   rgb MSE (:295-296), interlevel loss (:298-301), distortion loss (:303-305, the
//   inherited metrics_dict["distortion"]), thermal MSE (:319-324).
// The loss arithmetic itself is nerfstudio.model_components.losses:
//   lossfun_outer / interlevel_loss (searchsorted outer measure) and the mip-NeRF-360
//   distortion loss, here in its O(S) prefix-sum form.
// One warp per ray, grid-stride; every loss is a mean over rays (x samples), so each ray's
// gradient is local and forward + backward fuse into one pass.
#include "tnf_device.cuh"
#include "tnf_host.h"

namespace tnf {

constexpr float kLossEps = 1.0e-7f;

struct LossScratch {
  float c[kMaxFieldS + 8];   // final-level spacing bins (S2+1)
  float w[kMaxFieldS + 8];   // final-level weights
  float wi[kMaxFieldS + 8];  // inclusive prefix of w
  float wm[kMaxFieldS + 8];  // inclusive prefix of w*m
  float tp[kBuf];            // proposal bins (S_k+1)
  float cy[kBuf];            // cy1: exclusive prefix of proposal weights (S_k+1)
  float diff[kBuf];          // range-add difference array for d/d wp
};

// searchsorted(a[0..n), v, side="right"): number of elements <= v
__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads) tnf_losses_kernel(const __grid_constant__ TnfLossArgs a) {
  __shared__ LossScratch scratch[kWarpsPerCta];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LossScratch& s = scratch[warp];
  const long long R = a.num_rays;
  const int S2 = a.num_samples[TNF_NUM_PROP];
  const float invR = 1.f / (float)R;
  float l_rgb = 0.f, l_inter = 0.f, l_dist = 0.f, l_th = 0.f;

  for (long long ray = (long long)blockIdx.x * kWarpsPerCta + warp; ray < R;
       ray += (long long)gridDim.x * kWarpsPerCta) {
    // ---- MSE terms
    if (a.concat) {
      // ConcatNerfModel: MSE over the 4 RGBT channels of pred + noise (1 - accumulation)
      // (rgbt_renderer.py:134-140 blends the random background into the prediction only)
      float d = 0.f, n = 0.f;
      if (lane < 4) {
        const float one_minus_acc = 1.f - a.accumulation[ray];
        const float p = lane < 3 ? a.rgb[ray * 3 + lane] : a.thermal[ray];
        const float gt = lane < 3 ? a.gt_rgb[ray * 3 + lane] : a.gt_thermal[ray];
        n = a.noise[ray * 4 + lane];
        d = p + n * one_minus_acc - gt;
        l_rgb += d * d;
        const float g = a.grad_scale * 2.f * d * invR * 0.25f;
        if (lane < 3) { if (a.g_rgb) a.g_rgb[ray * 3 + lane] = g; }
        else if (a.g_thermal) a.g_thermal[ray] = g;
      }
      float ga = -(a.grad_scale * 2.f * d * invR * 0.25f) * n;  // d/d accumulation
      ga += __shfl_xor_sync(kFull, ga, 1);
      ga += __shfl_xor_sync(kFull, ga, 2);
      if (lane == 0 && a.g_accumulation) a.g_accumulation[ray] = ga;
    } else if (a.use_rgb_loss && lane < 3) {
      const float d = a.rgb[ray * 3 + lane] - a.gt_rgb[ray * 3 + lane];
      l_rgb += d * d;
      if (a.g_rgb) a.g_rgb[ray * 3 + lane] = a.grad_scale * 2.f * d * invR * (1.f / 3.f);
    } else if (!a.use_rgb_loss && lane < 3 && a.g_rgb) {
      a.g_rgb[ray * 3 + lane] = 0.f;
    }
    if (lane == 0 && !a.concat) {
      if (a.use_thermal_loss) {
        const float d = a.thermal[ray] - a.gt_thermal[ray];
        l_th += d * d;
        if (a.g_thermal) a.g_thermal[ray] = a.grad_scale * 2.f * d * invR;
      } else if (a.g_thermal) {
        a.g_thermal[ray] = 0.f;
      }
    }

    // ---- final level into shared memory
    for (int i = lane; i <= S2; i += 32) s.c[i] = a.sdist[TNF_NUM_PROP][ray * (S2 + 1) + i];
    for (int i = lane; i < S2; i += 32) s.w[i] = a.weights[TNF_NUM_PROP][ray * S2 + i];
    __syncwarp();

    // ---- distortion loss: sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 (c_{i+1} - c_i)
    {
      float cw = 0.f, cwm = 0.f;
      for (int base = 0; base < S2; base += 32) {
        const int i = base + lane;
        const float w = i < S2 ? s.w[i] : 0.f;
        const float m = i < S2 ? 0.5f * (s.c[i] + s.c[i + 1]) : 0.f;
        const float iw = warp_incl_scan(w, lane) + cw;
        const float iwm = warp_incl_scan(w * m, lane) + cwm;
        if (i < S2) { s.wi[i] = iw; s.wm[i] = iwm; }
        cw = __shfl_sync(kFull, iw, 31);
        cwm = __shfl_sync(kFull, iwm, 31);
      }
      __syncwarp();
      const float wtot = cw, wmtot = cwm;
      float part = 0.f;
      for (int i = lane; i < S2; i += 32) {
        const float w = s.w[i], m = 0.5f * (s.c[i] + s.c[i + 1]), dlt = s.c[i + 1] - s.c[i];
        const float wlt = s.wi[i] - w, wmlt = s.wm[i] - w * m;            // strictly before i
        const float wgt = wtot - s.wi[i], wmgt = wmtot - s.wm[i];         // strictly after i
        const float inner = m * wlt - wmlt + wmgt - m * wgt;              // sum_j w_j |m_i - m_j|
        part += w * inner + w * w * dlt * (1.f / 3.f);
        if (a.g_weights[TNF_NUM_PROP])
          a.g_weights[TNF_NUM_PROP][ray * S2 + i] =
              a.grad_scale * a.distortion_mult * invR * (2.f * inner + (2.f / 3.f) * w * dlt);
      }
      l_dist += part;
    }

    // ---- interlevel loss against each proposal level (final weights/bins are detached)
    for (int k = 0; k < TNF_NUM_PROP; ++k) {
      const int Sk = a.num_samples[k];
      for (int i = lane; i <= Sk; i += 32) {
        s.tp[i] = a.sdist[k][ray * (Sk + 1) + i];
        s.diff[i] = 0.f;
      }
      float carry = 0.f;
      if (lane == 0) s.cy[0] = 0.f;
      for (int base = 0; base < Sk; base += 32) {
        const int i = base + lane;
        const float wp = i < Sk ? a.weights[k][ray * Sk + i] : 0.f;
        const float inc = warp_incl_scan(wp, lane) + carry;
        if (i < Sk) s.cy[i + 1] = inc;
        carry = __shfl_sync(kFull, inc, 31);
      }
      __syncwarp();
      const float gnorm = a.grad_scale * a.interlevel_mult * invR / (float)S2;
      float part = 0.f;
      for (int i = lane; i < S2; i += 32) {
        int lo = upper_bound(s.tp, Sk, s.c[i]) - 1;           // over t_env[:-1]
        int hi = upper_bound(s.tp + 1, Sk, s.c[i + 1]);       // over t_env[1:]
        lo = min(max(lo, 0), Sk - 1);
        hi = min(max(hi, 0), Sk - 1);
        const float w_outer = s.cy[hi + 1] - s.cy[lo];
        const float w = s.w[i];
        const float d = fmaxf(w - w_outer, 0.f);
        part += d * d / (w + kLossEps);
        if (a.g_weights[k] && d > 0.f) {
          const float coef = -2.f * d / (w + kLossEps) * gnorm;
          atomicAdd(&s.diff[lo], coef);
          atomicAdd(&s.diff[hi + 1], -coef);
        }
      }
      l_inter += part / (float)S2;
      __syncwarp();
      if (a.g_weights[k]) {
        float run = 0.f;
        for (int base = 0; base < Sk; base += 32) {
          const int j = base + lane;
          const float v = j < Sk ? s.diff[j] : 0.f;
          const float inc = warp_incl_scan(v, lane) + run;
          if (j < Sk) a.g_weights[k][ray * Sk + j] = inc;
          run = __shfl_sync(kFull, inc, 31);
        }
      }
      __syncwarp();
    }
  }
  l_rgb = warp_sum(l_rgb);
  l_inter = warp_sum(l_inter);
  l_dist = warp_sum(l_dist);
  l_th = warp_sum(l_th);
  if (lane == 0) {
    if (a.concat) atomicAdd(a.losses + 0, l_rgb * invR * 0.25f);
    else if (a.use_rgb_loss) atomicAdd(a.losses + 0, l_rgb * invR * (1.f / 3.f));
    atomicAdd(a.losses + 1, a.interlevel_mult * l_inter * invR);
    atomicAdd(a.losses + 2, a.distortion_mult * l_dist * invR);
    if (a.use_thermal_loss && !a.concat) atomicAdd(a.losses + 3, l_th * invR);
  }
}

}  // namespace tnf

extern "C" int tnf_losses(const TnfLossArgs* args, void* stream_) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!args) return fail(TNF_ERR_INVALID_ARGUMENT, "args is null");
  const TnfLossArgs& a = *args;
  if (a.num_rays < 0) return fail(TNF_ERR_INVALID_ARGUMENT, "num_rays=%lld", (long long)a.num_rays);
  if (!a.losses) return fail(TNF_ERR_INVALID_ARGUMENT, "losses is null");
  for (int k = 0; k <= TNF_NUM_PROP; ++k) {
    if (!a.weights[k] || !a.sdist[k]) return fail(TNF_ERR_INVALID_ARGUMENT, "weights[%d]/sdist[%d] is null", k, k);
    const int lim = (k == TNF_NUM_PROP) ? tnf::kMaxFieldS : TNF_MAX_SAMPLES;
    if (a.num_samples[k] < 1 || a.num_samples[k] > lim)
      return fail(TNF_ERR_UNSUPPORTED_CONFIG, "num_samples[%d]=%d not in [1,%d]", k, a.num_samples[k], lim);
  }
  if (a.concat && (!a.rgb || !a.gt_rgb || !a.thermal || !a.gt_thermal || !a.accumulation || !a.noise))
    return fail(TNF_ERR_INVALID_ARGUMENT, "concat loss needs rgb, thermal, gt_rgb, gt_thermal, accumulation and noise");
  if (a.use_rgb_loss && (!a.rgb || !a.gt_rgb)) return fail(TNF_ERR_INVALID_ARGUMENT, "rgb/gt_rgb is null");
  if (a.use_thermal_loss && !a.concat && (!a.thermal || !a.gt_thermal))
    return fail(TNF_ERR_INVALID_ARGUMENT, "thermal/gt_thermal is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(a.losses, 0, 4 * sizeof(float), stream);
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  if (a.num_rays == 0) return TNF_OK;
  const long long want = (a.num_rays + tnf::kWarpsPerCta - 1) / tnf::kWarpsPerCta;
  const long long cap = (long long)tnf::num_sms() * 4;
  tnf::tnf_losses_kernel<<<(unsigned)(want < cap ? want : cap), tnf::kThreads, 0, stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "losses kernel launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}
