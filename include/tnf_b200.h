/*
 * tnf_b200.h - C ABI of libtnf_b200.so: the ThermoNeRF volumetric-render hot path
 * as hand-written sm_100a CUDA.
 *
 * Boundary contract:
 *   - plain pointers and sizes only; no torch types.  Every buffer is owned by the
 *     caller (a PyTorch CUDA tensor in the Python host layer): contiguous, fp32
 *     (int64 for camera indices), 16-byte aligned.  The library never allocates or
 *     frees caller-visible memory and keeps no state between calls except a
 *     thread-local error string.
 *   - every entry point returns TNF_OK (0) or a negative TnfStatus and never throws;
 *     tnf_last_error() describes the last failure on the calling thread.
 *   - all work is enqueued on the cudaStream_t passed as `stream` (opaque void*);
 *     nothing synchronises the device.
 *
 * Reference interface each entry point replaces (paths relative to the reference
 * repo Schindler-EPFL-Lab/thermo-nerf; the arithmetic itself lives in
 * nerfstudio==1.1.5, reference pyproject.toml:11):
 *
 *   tnf_render_forward   <- ThermalNerfModel.get_outputs
 *                           thermo_nerf/thermal_nerf/thermal_nerf_model.py:210-275
 *                           (proposal sampler :222-224, field.forward :225-227 =
 *                           thermal_field.py:108-201, get_weights :233, renderers
 *                           :237-243, prop depths :267-270, ThermalRenderer :271 =
 *                           thermal_renderer.py:113-149), plus the NearFarCollider
 *                           built at thermal_nerf_model.py:182-184.
 *   tnf_render_backward  <- autograd of the same function as driven by
 *                           get_loss_dict, thermal_nerf_model.py:277-326.
 *   tnf_losses           <- ThermalNerfModel.get_loss_dict, thermal_nerf_model.py:277-326
 *                           (rgb MSE :295-296, interlevel :298-301, distortion :303-305,
 *                           thermal MSE :319-324).
 *   tnf_adam_step        <- the Adam optimisers configured at
 *                           thermo_nerf/thermal_nerf/config_thermal_nerf.py:32-45.
 */
#ifndef TNF_B200_H_
#define TNF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNF_ABI_VERSION 9

#define TNF_MAX_LEVELS 16      /* hash levels of the field grid (fixed: 16)          */
#define TNF_MAX_PROP_LEVELS 8  /* max hash levels of a proposal density grid          */
#define TNF_MAX_SAMPLES 256    /* max samples per ray of any level                    */
#define TNF_NUM_PROP 2         /* proposal iterations (nerfacto default, fixed)       */

typedef enum TnfStatus {
  TNF_OK = 0,
  TNF_ERR_INVALID_ARGUMENT = -1,   /* null / misaligned pointer, negative size       */
  TNF_ERR_UNSUPPORTED_CONFIG = -2, /* architecture outside what the kernels compile  */
  TNF_ERR_WORKSPACE_TOO_SMALL = -3,
  TNF_ERR_CUDA = -4                /* a CUDA runtime call failed; see tnf_last_error */
} TnfStatus;

/* How the 32-d appearance embedding is chosen (thermal_field.py:124-137). */
typedef enum TnfAppearanceMode {
  TNF_APPEARANCE_ZEROS = 0,  /* eval, use_average_appearance_embedding=False        */
  TNF_APPEARANCE_MEAN = 1,   /* eval, mean over all training cameras                 */
  TNF_APPEARANCE_LOOKUP = 2  /* training: embedding[camera_indices]                  */
} TnfAppearanceMode;

/* Arithmetic used for the 64-wide field MLPs. Hashing, sampling, the proposal MLPs
 * and compositing are fp32 in both modes. */
typedef enum TnfPrecision {
  TNF_PRECISION_FP32 = 0,    /* fp32 FFMA everywhere                                  */
  TNF_PRECISION_TC_FP16 = 1  /* fp16 operands, fp32 accumulate on the tensor cores    */
} TnfPrecision;

/* nerfstudio HashEncoding, torch layout: table [num_levels << log2_size, 2] fp32,
 * level-major; `scalings` is the registered buffer (host copy). */
/* Which field head produces the temperature channel.
 *   THERMAL  ThermalNerfModel: colour head 63-64-64-3 + thermal head 15-64-64-1, both composited with the
 *            last sample as background (thermal_field.py:160-179, thermal_renderer.py:49).
 *   CONCAT   ConcatNerfModel (the `concat_nerf` model type, train_eval_script.py:74-78): one 4-channel RGBT
 *            colour head (rgb_concat/concat_field.py:65-75: field.rgb2 is [4,64] / [4]), no thermal head (its
 *            tensors are ignored), RGBTRenderer with its default "random" background = plain weighted sum, no
 *            background term (rgb_concat/rgbt_renderer.py:63-71).  out.rgb carries channels 0-2 and out.thermal
 *            channel 3 of the 4-channel "rgb" output. */
typedef enum TnfHeadMode {
  TNF_HEAD_THERMAL = 0,
  TNF_HEAD_CONCAT = 1
} TnfHeadMode;

typedef struct TnfHashGrid {
  const float* table;
  float scalings[TNF_MAX_LEVELS];
  int32_t num_levels;
  int32_t log2_size;
} TnfHashGrid;

/* torch.nn.Linear: weight [out, in] row-major, bias [out]. */
typedef struct TnfLinear {
  const float* weight;
  const float* bias;
} TnfLinear;

/* HashMLPDensityField (thermal_nerf_model.py:127-148): grid -> 16 -> 1. */
typedef struct TnfDensityNet {
  TnfHashGrid grid;
  TnfLinear l0; /* [16, 2*num_levels] */
  TnfLinear l1; /* [1, 16]            */
} TnfDensityNet;

/* ThermalNerfactoTField (thermal_field.py:33-106). */
typedef struct TnfField {
  TnfHashGrid grid;        /* 16 levels x 2 features                              */
  TnfLinear base0;         /* mlp_base.mlp.layers.0        [64, 32]               */
  TnfLinear base1;         /* mlp_base.mlp.layers.1        [16, 64] (1 | 15 geo)  */
  TnfLinear rgb0;          /* mlp_head.layers.0            [64, 63] sh|geo|app    */
  TnfLinear rgb1;          /* mlp_head.layers.1            [64, 64]               */
  TnfLinear rgb2;          /* mlp_head.layers.2            [3, 64]                */
  TnfLinear th0;           /* mlp_thermal.layers.0         [64, 15]               */
  TnfLinear th1;           /* mlp_thermal.layers.1         [64, 64] (sigmoid out) */
  TnfLinear th2;           /* field_head_thermal.net       [1, 64]                */
  const float* appearance; /* embedding_appearance weight  [num_images, 32]       */
  int32_t num_images;
  int32_t _pad;
} TnfField;

typedef struct TnfModel {
  TnfDensityNet prop[TNF_NUM_PROP];
  TnfField field;
  int32_t num_samples[TNF_NUM_PROP + 1]; /* (256, 96, 48) by default; each <= 256, last <= 64 */
  int32_t training;        /* 0: eval renderers (nan_to_num + clamp[0,1]); 1: training        */
  float near_plane;        /* used when rays.nears == NULL                                    */
  float far_plane;         /* used when rays.fars  == NULL                                    */
  float anneal;            /* ProposalNetworkSampler._anneal (1.0 for a fresh/eval sampler)   */
  int32_t use_contraction; /* 1: L-inf scene contraction; 0: aabb normalisation               */
  float aabb[6];           /* min xyz, max xyz (only for use_contraction == 0)                */
  int32_t appearance_mode; /* TnfAppearanceMode                                               */
  int32_t precision;       /* TnfPrecision                                                    */
  int32_t detach_thermal_geo; /* 1: pass_thermal_gradients == False (thermal_field.py:173-175): the
                                 thermal head's gradient stops at the geo feature                */
  int32_t head_mode;       /* TnfHeadMode                                                      */
} TnfModel;

/* One perspective camera without distortion, as nerfstudio's Cameras holds it (the cameras the reference
 * renders: thermo_nerf/render/renderer.py:144-157 loads them, :183 and evaluator.py:69 call generate_rays). */
typedef struct TnfCamera {
  float c2w[12];          /* camera_to_worlds[i], row-major [3,4]                                  */
  float fx, fy, cx, cy;
  int32_t width, height;
} TnfCamera;

typedef struct TnfRays {
  const float* origins;          /* [R,3]                                                     */
  const float* directions;       /* [R,3]                                                     */
  const int64_t* camera_indices; /* [R]   (required for TNF_APPEARANCE_LOOKUP, else may be 0) */
  const float* nears;            /* [R] or NULL                                               */
  const float* fars;             /* [R] or NULL                                               */
  const float* jitter;           /* [3,R] single-jitter draws in [0,1) (training) or NULL     */
  int64_t num_rays;
  /* from_camera != 0 (tnf_render_forward, eval only): ray r is pixel first_pixel + r (row-major) of
   * `camera`, generated inside the kernel exactly as Cameras.generate_rays does (pixel centre +0.5,
   * ((x-cx)/fx, -(y-cy)/fy, -1) rotated by c2w and normalised); origins/directions may then be NULL. */
  int32_t from_camera;
  int32_t _pad;
  int64_t first_pixel;
  TnfCamera camera;
} TnfRays;

typedef struct TnfOutputs {
  float* rgb;            /* [R,3] */
  float* thermal;        /* [R,1] */
  float* depth;          /* [R,1] median depth                                              */
  float* expected_depth; /* [R,1] clipped to the per-chunk [min,max] of sample mid-points   */
  float* accumulation;   /* [R,1] */
  float* prop_depth[TNF_NUM_PROP]; /* [R,1] each                                            */
  /* optional (NULL to skip) - the training-only outputs of get_outputs: */
  float* weights[TNF_NUM_PROP + 1]; /* weights_list[k]           [R, S_k]                   */
  float* sdist[TNF_NUM_PROP + 1];   /* spacing bins of level k   [R, S_k + 1]               */
  /* optional (NULL to skip) - saved for tnf_render_backward: */
  void* field_features;  /* [R*S_2, 32] hash features of the field samples; fp32 for
                            TNF_PRECISION_FP32, fp16 for TNF_PRECISION_TC_FP16                 */
  float* field_samples;  /* [R*S_2, 5] per-sample field outputs (sigma, r, g, b, thermal)      */
} TnfOutputs;

/* ABI version of the loaded library (== TNF_ABI_VERSION of the header it was built from). */
int tnf_version(void);

/* Description of the last error on the calling thread ("" if none). */
const char* tnf_last_error(void);

/* Bytes of device scratch tnf_render_forward needs for `num_rays` rays when the
 * expected-depth clip is evaluated per `depth_clip_chunk` rays (<= 0: whole call). */
size_t tnf_forward_workspace_bytes(int64_t num_rays, int64_t depth_clip_chunk);

/*
 * One pass of get_outputs over `rays`.  `depth_clip_chunk` reproduces the
 * reference's chunk-dependent expected-depth clip (nerfstudio clips to the min/max
 * sample mid-point of each forward call, i.e. of each eval_num_rays_per_chunk slice,
 * config_thermal_nerf.py:30): rays [k*chunk, (k+1)*chunk) share one clip range.
 */
int tnf_render_forward(const TnfModel* model, const TnfRays* rays, const TnfOutputs* out,
                       int64_t depth_clip_chunk, void* workspace, size_t workspace_bytes,
                       void* stream);

/* tnf_render_forward whose field level additionally waits for `field_params_ready` (a cudaEvent_t, or NULL for no
 * wait) on `stream`: in training mode the call is two launches - the proposal levels (proposal-network parameters
 * only) and the field level + compositing - and the event is waited for between them, so a parameter exchange of the
 * field network (tnf_peer_adam_range on another stream) may still be in flight while the proposal pass runs.  When the
 * call is a single launch (eval, fp32 mode) the wait comes first. */
int tnf_render_forward_staged(const TnfModel* model, const TnfRays* rays, const TnfOutputs* out,
                              int64_t depth_clip_chunk, void* workspace, size_t workspace_bytes, void* stream,
                              void* field_params_ready);

/* Cameras.generate_rays for `num_pixels` pixels of one camera starting at first_pixel (row-major):
 * origins/directions [n,3], directions_norm [n] (the RayBundle metadata entry; may be NULL).
 * Replaces cameras.generate_rays(camera_indices=i) at thermo_nerf/render/renderer.py:183 and
 * thermo_nerf/evaluator/evaluator.py:69. */
int tnf_generate_rays(const TnfCamera* camera, int64_t first_pixel, int64_t num_pixels, float* origins,
                      float* directions, float* directions_norm, void* stream);

/* Frame post-processing of Renderer.render (thermo_nerf/render/renderer.py:189-199) on the device:
 *   rgb8[i,c]    = (uint8)(rgb[i,c] * 255)                    for an [n,3] float image in [0,1]
 *   scalar8[i,:] = lut8[index(scalar[i])]                     when lut8 != NULL (matplotlib Colormap.__call__:
 *                  index = scalar*lut_n truncated, scalar == 1 -> lut_n-1, clipped to the table; NaN -> 0,0,0)
 *                = (uint8)(scalar[i] * 255) replicated x3     when lut8 == NULL (single-channel modalities)
 * lut8 is the colour map already converted the reference's way, (cmap(arange(N))[:, :3] * 255).astype(uint8).
 * Either pair (rgb, rgb8) / (scalar, scalar8) may be NULL. */
int tnf_postprocess_frame(const float* rgb, const float* scalar, int64_t num_pixels, const uint8_t* lut8,
                          int32_t lut_n, uint8_t* rgb8, uint8_t* scalar8, void* stream);

/* ------------------------------------------------------------------------------------
 * Training: backward of tnf_render_forward, the losses of get_loss_dict, Adam.
 * ------------------------------------------------------------------------------------ */

/* Gradient buffers, same shapes as the parameters.  Every entry point ACCUMULATES (+=)
 * into them; the caller zeroes them (tnf_adam_step can do so after consuming them). */
typedef struct TnfLinearGrad {
  float* weight;
  float* bias;
} TnfLinearGrad;

typedef struct TnfDensityNetGrad {
  float* table; /* NULL: skip this proposal level entirely (the reference's no_grad steps,
                   ProposalNetworkSampler `updated == False`) */
  TnfLinearGrad l0;
  TnfLinearGrad l1;
} TnfDensityNetGrad;

typedef struct TnfFieldGrad {
  float* table;
  TnfLinearGrad base0, base1, rgb0, rgb1, rgb2, th0, th1, th2;
  float* appearance; /* [num_images, 32]; may be NULL unless TNF_APPEARANCE_LOOKUP */
} TnfFieldGrad;

typedef struct TnfModelGrad {
  TnfDensityNetGrad prop[TNF_NUM_PROP];
  TnfFieldGrad field;
  /* optional (both or neither; NULL to skip): dLoss/d origins and dLoss/d directions [R,3], accumulated
   * (+=).  They carry the gradient into the camera optimiser's pose deltas, which the reference applies to
   * the ray bundle before get_outputs (thermal_nerf_model.py:218-219).  Sample distances are constants
   * (the sampler detaches its bins), so x = o + t d gives dL/do = sum_s dL/dx_s, dL/dd = sum_s t_s dL/dx_s. */
  float* ray_origins;
  float* ray_directions;
} TnfModelGrad;

/* What tnf_render_forward wrote in training mode (same pointers as in TnfOutputs). */
typedef struct TnfSaved {
  const float* sdist[TNF_NUM_PROP + 1];
  const float* weights[TNF_NUM_PROP + 1];
  const void* field_features;
  const float* field_samples;
} TnfSaved;

/* dLoss/dOutput; any pointer may be NULL (= zero gradient). */
typedef struct TnfOutputGrads {
  const float* rgb;                     /* [R,3] */
  const float* thermal;                 /* [R]   */
  const float* accumulation;            /* [R]   */
  const float* weights[TNF_NUM_PROP + 1]; /* [R,S_k]: from the interlevel (k<2) and distortion (k=2) losses */
} TnfOutputGrads;

size_t tnf_backward_workspace_bytes(const TnfModel* model, int64_t num_rays);

/* Backward of tnf_render_forward w.r.t. every parameter (sample positions are detached
 * as in PDFSampler; depth outputs carry no gradient).  `model` and `rays` must be the
 * ones of the forward call (same jitter). */
int tnf_render_backward(const TnfModel* model, const TnfRays* rays, const TnfSaved* saved,
                        const TnfOutputGrads* gout, const TnfModelGrad* grads, void* workspace,
                        size_t workspace_bytes, void* stream);

/* tnf_render_backward with the two levels in the other order - field level first, then the proposal levels (their
 * gradients are disjoint: the interlevel loss sees the field's weights detached) - and `field_grads_done` (a
 * cudaEvent_t; NULL = plain tnf_render_backward) recorded on `stream` between them: a multi-GPU trainer starts the
 * exchange of the field gradients (tnf_peer_adam_range on another stream) while the proposal kernel still runs.
 * `reserve_ctas`: CTA slots of 256 threads x 64 registers the proposal kernel leaves free for that exchange kernel
 * (its grid shrinks; the work units are handed out by a device counter, so nothing else changes). */
int tnf_render_backward_staged(const TnfModel* model, const TnfRays* rays, const TnfSaved* saved,
                               const TnfOutputGrads* gout, const TnfModelGrad* grads, void* workspace,
                               size_t workspace_bytes, void* stream, void* field_grads_done, int32_t reserve_ctas);

/* ---- Field / Renderer plugin surface: the per-module entry points of the reference's
 * ThermalNerfactoTField and ThermalRenderer for callers that compose the modules by hand (viewer density
 * queries, user code).  fp32, inference only; training and rendering go through tnf_render_forward. */

/* NerfactoField.get_density as reached from ThermalNerfactoTField.forward (thermal_field.py:183-201)
 * for which = -1 (density [N] and, if `geo` is given, the 15 geo features [N,15]); or
 * HashMLPDensityField.density_fn of proposal network which = 0 / 1 (thermal_nerf_model.py:127-148).
 * positions [N,3] are world-space points (contraction + selector applied inside, as in the field). */
int tnf_field_density(const TnfModel* model, int32_t which, const float* positions, int64_t n, float* density,
                      float* geo, void* stream);

/* ThermalNerfactoTField.get_outputs (thermal_field.py:108-181): directions [N,3], camera_indices [N]
 * (TNF_APPEARANCE_LOOKUP only), density_embedding geo [N,15] -> rgb [N,3] (FieldHeadNames.RGB) and
 * thermal [N] (FieldHeadNamesT.THERMAL); either output may be NULL. */
int tnf_field_heads(const TnfModel* model, const float* directions, const int64_t* camera_indices, const float* geo,
                    int64_t n, float* rgb, float* thermal, void* stream);

/* ThermalRenderer.forward (thermal_renderer.py:113-149; last-sample background, :49,68-70) for
 * last_sample_background = 1, RGBTRenderer.forward of the concat baseline (rgb_concat/rgbt_renderer.py:134-140)
 * for 0: values [R,S,C], weights [R,S] -> out [R,C]; eval_mode applies nan_to_num to the samples and
 * clamps the result to [0,1] (thermal_renderer.py:136-137,146-147). */
int tnf_composite(const float* values, const float* weights, int64_t num_rays, int32_t num_samples,
                  int32_t channels, int32_t last_sample_background, int32_t eval_mode, float* out, void* stream);

/* Measurement hook (bench.py's per-kernel roofline): restricts tnf_render_backward on the calling thread
 * to a subset of its kernels so each can be bracketed with CUDA events on its own.  Bit 0: proposal
 * levels, bit 1: field level, bit 2: the fp32 mode's weight-gradient pass (tensor-core mode accumulates
 * the weight gradients inside the field kernel); the default 7 runs all of them (the only setting that
 * produces correct gradients).  Returns the previous mask. */
int tnf_backward_stage_mask(int mask);

/* get_loss_dict (thermal_nerf_model.py:277-326) + the inherited distortion metric:
 *   losses[0] rgb_loss        = MSE(gt_rgb, rgb)                       (if use_rgb_loss)
 *   losses[1] interlevel_loss = interlevel_mult * sum_k mean(outer-measure loss of level k)
 *   losses[2] distortion_loss = distortion_mult * mean_r(distortion of the final level)
 *   losses[3] thermal         = MSE(thermal, gt_thermal)               (if use_thermal_loss)
 * and, when the g_* pointers are non-NULL, d(loss_i)/d(input) * grad_scale for the one
 * loss each input feeds (rgb <- 0, weights[0..1] <- 1, weights[2] <- 2, thermal <- 3).
 * `losses` (4 floats, device) is overwritten. */
typedef struct TnfLossArgs {
  const float* weights[TNF_NUM_PROP + 1];
  const float* sdist[TNF_NUM_PROP + 1];
  const float* rgb;        /* [R,3] */
  const float* thermal;    /* [R]   */
  const float* gt_rgb;     /* [R,3] */
  const float* gt_thermal; /* [R]   */
  int64_t num_rays;
  int32_t num_samples[TNF_NUM_PROP + 1];
  float interlevel_mult;
  float distortion_mult;
  int32_t use_rgb_loss;
  int32_t use_thermal_loss;
  float grad_scale;
  float* losses;                       /* [4] */
  float* g_rgb;                        /* [R,3] or NULL */
  float* g_thermal;                    /* [R]   or NULL */
  float* g_weights[TNF_NUM_PROP + 1];  /* [R,S_k] or NULL */
  /* concat_nerf loss (rgb_concat/concat_nerfacto_model.py:197-233, rgbt_renderer.py:118-140): with concat != 0
   * losses[0] = MSE over the 4 channels [rgb | thermal] of pred + noise * (1 - accumulation) against
   * [gt_rgb | gt_thermal] (the renderer's "random" background is blended into the prediction only; `noise` is the
   * caller's torch.rand_like(pred) draw), losses[3] = 0, and g_accumulation receives d losses[0] / d accumulation. */
  int32_t concat;
  int32_t _pad;
  const float* accumulation;  /* [R]   */
  const float* noise;         /* [R,4] */
  float* g_accumulation;      /* [R] or NULL */
} TnfLossArgs;

int tnf_losses(const TnfLossArgs* args, void* stream);

/* torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False) over a list of tensors
 * in ONE launch.  grad is first multiplied by inv_grad_scale (host value) and, if grad_scale
 * (device float, may be NULL) is given, divided by *grad_scale; if found_inf (device float,
 * may be NULL) is non-zero the update is skipped (torch.amp.GradScaler protocol of optimisers
 * with _step_supports_amp_scaling); with zero_grads != 0 the gradients are zeroed after use
 * (also on a skipped step). */
#define TNF_ADAM_MAX_TENSORS 48
typedef struct TnfAdamTensor {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  float lr;
  int32_t _pad;
} TnfAdamTensor;

int tnf_adam_step(const TnfAdamTensor* tensors, int32_t num_tensors, double beta1, double beta2, float eps,
                  int64_t step, float inv_grad_scale, const float* grad_scale, const float* found_inf,
                  int32_t zero_grads, void* stream);

/* ------------------------------------------------------------------------------------
 * Training batches from a device-resident dataset: nerfstudio's PixelSampler.sample_method +
 * collate_image_dataset_batch + RayGenerator (VanillaDataManager.next_train, driven from
 * thermo_nerf/nerfstudio_config/pipeline_tracking.py:47-59) in one launch.  The reference keeps the
 * thermal ground truth on the host and moves a batch per step (thermal_dataset.py:18-20,
 * thermal_nerf_model.py:319).
 * ------------------------------------------------------------------------------------ */
typedef struct TnfDataset {
  const void* images;            /* [N,H,W,C] float32 in [0,1], or uint8 (converted as x / 255)        */
  const void* thermal;           /* [N,H,W]   float32 normalised temperature, or uint8; may be NULL   */
  const float* camera_to_worlds; /* [N,3,4]                                                            */
  const float* intrinsics;       /* [N,4] fx, fy, cx, cy                                               */
  int32_t num_images, height, width, channels; /* channels >= 3 (the first three are RGB)              */
  int32_t images_uint8, thermal_uint8;
} TnfDataset;

/* indices = floor(rand * (N, H, W)) per ray; gathers the ground truth at those pixels and generates the
 * rays through their centres.  rand [R,3] uniform [0,1) (the caller draws it, e.g. torch.rand on the device);
 * outputs: origins / directions [R,3], camera_indices [R] int64, indices [R,3] int64 (image, y, x; may be
 * NULL), gt_rgb [R,3] and gt_thermal [R] (each may be NULL). */
int tnf_sample_batch(const TnfDataset* dataset, const float* rand, int64_t num_rays, float* origins,
                     float* directions, int64_t* camera_indices, int64_t* indices, float* gt_rgb, float* gt_thermal,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU exchange step (one process per GPU): gradient mean over ranks fused with Adam
 * over NVLink peer memory.  Replaces the DDP gradient all-reduce of nerfstudio's trainer followed by the
 * optimizers of config_thermal_nerf.py:32-45 (reduce-scatter -> Adam on the owned shard -> all-gather in
 * one kernel).  All three buffers of every rank must be mapped into this process (CUDA IPC) and peer
 * access enabled from the current device (tnf_peer_enable_access).
 * ------------------------------------------------------------------------------------ */
#define TNF_MAX_PEERS 16
#define TNF_PEER_FLAG_SLOTS 4       /* independent barrier slots (TNF_MAX_PEERS words each)        */
#define TNF_PEER_FLAG_TIMEOUT 64    /* word of a rank's own flag block that counts timed-out waits */
#define TNF_PEER_FLAG_WORDS 128     /* uint32 words per flag block (zero-initialised by the owner) */
#define TNF_MAX_ADAM_SEGMENTS 4

typedef struct TnfPeerArena {
  const float* grads[TNF_MAX_PEERS]; /* flat gradient arena of every rank, `numel` floats          */
  float* params[TNF_MAX_PEERS];      /* flat parameter arena of every rank                          */
  uint32_t* flags[TNF_MAX_PEERS];    /* flag block of every rank, TNF_PEER_FLAG_WORDS uint32        */
  int32_t world_size;
  int32_t rank;
  int64_t numel;                     /* multiple of 4 * world_size; rank r owns [r, r+1) * numel/N  */
} TnfPeerArena;

/* A run of the arena that shares one optimizer group (learning rate, step counter); inactive segments are
 * left untouched (a param group without gradients this step, i.e. the proposal networks on the sampler's
 * no_grad steps).  Segments are contiguous, start at 0 and cover the arena. */
typedef struct TnfAdamSegment {
  int64_t begin, end;
  int64_t step;   /* >= 1 when active */
  float lr;
  int32_t active;
} TnfAdamSegment;

/* cudaDeviceEnablePeerAccess(peer_device) from the current device (no-op for the device itself). */
int tnf_peer_enable_access(int32_t peer_device);

/* Map a buffer another process exported with cudaIpcGetMemHandle (64-byte handle) for the CURRENT device
 * (cudaIpcOpenMemHandle + lazy peer access), and unmap it again.  *device_ptr is the base of the exported
 * allocation. */
#define TNF_IPC_HANDLE_BYTES 64
/* A zero-filled device allocation of its own (cudaMalloc) on the current device plus the 64-byte
 * cudaIpcMemHandle_t other processes map it with; tnf_peer_free releases it. */
int tnf_peer_alloc(size_t bytes, void** device_ptr, void* ipc_handle);
int tnf_peer_free(void* device_ptr);
int tnf_peer_open_handle(const void* ipc_handle, void** device_ptr);
int tnf_peer_close_handle(void* device_ptr);

/* Flag barrier across the ranks of `arena` on `stream` (slot in [0, TNF_PEER_FLAG_SLOTS), epoch strictly
 * increasing per slot, same value on every rank). */
int tnf_peer_barrier(const TnfPeerArena* arena, int32_t slot, uint32_t epoch, void* stream);

/* The fused reduce-scatter(mean) + Adam + all-gather described above.  exp_avg / exp_avg_sq hold the
 * state of this rank's shard only (numel / world_size floats each).  Bracket with tnf_peer_barrier:
 * before (every rank's gradients are complete) and after (every rank's parameters are written). */
int tnf_peer_adam_step(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                       const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                       void* stream);

/* The same exchange with a pull-style all-gather, as two launches: tnf_peer_adam_reduce (reduce-scatter + Adam,
 * the updated shard is stored into this rank's own parameters only), a tnf_peer_barrier, then
 * tnf_peer_gather_params (every rank copies the other ranks' shards out of their owners' arenas with peer
 * loads).  Same results bit for bit; which flavour is faster depends on the fabric. */
int tnf_peer_adam_reduce(const TnfPeerArena* arena, float* exp_avg_shard, float* exp_avg_sq_shard,
                         const TnfAdamSegment* segments, int32_t num_segments, double beta1, double beta2, float eps,
                         void* stream);
/* The same exchange through NVSwitch multicast (NVLS): `grads_multicast` / `params_multicast` are the multicast
 * addresses of the gradient and parameter arenas (every rank's arena bound to one multicast object, e.g. by
 * torch.distributed._symmetric_memory).  multimem.ld_reduce forms the sum over ranks inside the switch and
 * multimem.st replicates the updated shard into every rank's parameters: each rank moves 1/N of the arena per
 * direction instead of (N-1)/N.  Replaces DDP's all-reduce + the optimisers of config_thermal_nerf.py:32-45 like
 * tnf_peer_adam_step; bracket it with the same two tnf_peer_barrier calls. */
int tnf_peer_adam_multimem(const TnfPeerArena* arena, const float* grads_multicast, float* params_multicast,
                           float* exp_avg_shard, float* exp_avg_sq_shard, const TnfAdamSegment* segments,
                           int32_t num_segments, double beta1, double beta2, float eps, void* stream);

int tnf_peer_gather_params(const TnfPeerArena* arena, const TnfAdamSegment* segments, int32_t num_segments,
                           void* stream);

/* The fused exchange over the slice [range_begin, range_end) of the arenas only, sharded over the ranks on its own
 * (rank r owns the r-th of world_size equal parts of the slice; the slice's length must be a multiple of
 * 4 * world_size).  exp_avg / exp_avg_sq hold the state of this rank's part of THIS slice.  `flavour`:
 * TNF_PEER_PUSH = the kernel of tnf_peer_adam_step, TNF_PEER_MULTIMEM = the kernel of tnf_peer_adam_multimem (needs
 * the two multicast addresses, ignored otherwise).  `max_ctas` > 0 bounds the grid so that the exchange can share
 * the SMs with a compute kernel running beside it on another stream.
 * Why slices: the proposal networks and the field are separate optimizer groups of config_thermal_nerf.py:32-45 and
 * the next iteration's proposal pass reads only the former, so the trainer exchanges the proposal slice first and
 * lets the (seven times larger) field slice travel while that pass already runs; see tnf_render_forward_staged. */
#define TNF_PEER_PUSH 0
#define TNF_PEER_MULTIMEM 3
int tnf_peer_adam_range(const TnfPeerArena* arena, int32_t flavour, int64_t range_begin, int64_t range_end,
                        int32_t max_ctas, const float* grads_multicast, float* params_multicast,
                        float* exp_avg_shard, float* exp_avg_sq_shard, const TnfAdamSegment* segments,
                        int32_t num_segments, double beta1, double beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNF_B200_H_ */
