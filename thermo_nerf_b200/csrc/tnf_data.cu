// Training-side neighbour of the path: nerfstudio's PixelSampler + collate + RayGenerator
// (what VanillaDataManager.next_train does before ThermalNerfModel.get_outputs, reached from
// thermo_nerf/nerfstudio_config/pipeline_tracking.py:47-59) as one kernel over a device-resident dataset:
//
//   indices = floor(rand[R,3] * (num_images, H, W))                    PixelSampler.sample_method
//   gt_rgb  = images[c, y, x, :3]   gt_thermal = thermal[c, y, x]      collate_image_dataset_batch
//   rays    = cameras.generate_rays(c, coords = (y + 0.5, x + 0.5))    RayGenerator.forward
//
// The reference keeps the thermal images on the host and moves a batch per step (thermal_dataset.py:18-20,
// thermal_nerf_model.py:319); here both modalities stay in HBM and a step moves nothing across PCIe.
#include "tnf_field.cuh"
#include "tnf_host.h"

namespace tnf {

struct SampleArgs {
  const float* rand;           // [R,3] uniform [0,1)
  const void* images;          // [N,H,W,C] float32 or uint8
  const void* thermal;         // [N,H,W] float32 or uint8 (may be null)
  const float* c2w;            // [N,3,4]
  const float* intrinsics;     // [N,4] fx, fy, cx, cy
  long long R;
  int N, H, W, C;
  int images_u8, thermal_u8;
  float* origins; float* directions; long long* camera_indices; long long* indices;  // [R,3] [R,3] [R] [R,3]
  float* gt_rgb; float* gt_thermal;                                                    // [R,3] [R]
};

__global__ void tnf_sample_batch_kernel(const __grid_constant__ SampleArgs a) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  // torch.floor(rand * tensor([N, H, W])).long(): float32 product, floor, truncate
  const long long c = (long long)floorf(__fmul_rn(a.rand[3 * r + 0], (float)a.N));
  const long long y = (long long)floorf(__fmul_rn(a.rand[3 * r + 1], (float)a.H));
  const long long x = (long long)floorf(__fmul_rn(a.rand[3 * r + 2], (float)a.W));
  if (a.indices) { a.indices[3 * r] = c; a.indices[3 * r + 1] = y; a.indices[3 * r + 2] = x; }
  a.camera_indices[r] = c;
  const long long pix = (c * a.H + y) * a.W + x;
  if (a.gt_rgb) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a.gt_rgb[3 * r + k] = a.images_u8
          ? (float)static_cast<const unsigned char*>(a.images)[pix * a.C + k] / 255.0f
          : static_cast<const float*>(a.images)[pix * a.C + k];
    }
  }
  if (a.gt_thermal && a.thermal) {
    a.gt_thermal[r] = a.thermal_u8 ? (float)static_cast<const unsigned char*>(a.thermal)[pix] / 255.0f
                                   : static_cast<const float*>(a.thermal)[pix];
  }
  TnfCamera cam;
#pragma unroll
  for (int k = 0; k < 12; ++k) cam.c2w[k] = a.c2w[c * 12 + k];
  cam.fx = a.intrinsics[c * 4 + 0];
  cam.fy = a.intrinsics[c * 4 + 1];
  cam.cx = a.intrinsics[c * 4 + 2];
  cam.cy = a.intrinsics[c * 4 + 3];
  cam.width = a.W;
  cam.height = a.H;
  float ox, oy, oz, dx, dy, dz, nrm;
  camera_ray(cam, y * a.W + x, ox, oy, oz, dx, dy, dz, nrm);
  a.origins[3 * r] = ox; a.origins[3 * r + 1] = oy; a.origins[3 * r + 2] = oz;
  a.directions[3 * r] = dx; a.directions[3 * r + 1] = dy; a.directions[3 * r + 2] = dz;
}

}  // namespace tnf

extern "C" int tnf_sample_batch(const TnfDataset* ds, const float* rand, int64_t num_rays, float* origins,
                                float* directions, int64_t* camera_indices, int64_t* indices, float* gt_rgb,
                                float* gt_thermal, void* stream_) {
  using tnf::fail;
  tnf::g_err[0] = 0;
  if (!ds) return fail(TNF_ERR_INVALID_ARGUMENT, "dataset is null");
  if (num_rays < 0) return fail(TNF_ERR_INVALID_ARGUMENT, "num_rays=%lld", (long long)num_rays);
  if (ds->num_images < 1 || ds->height < 1 || ds->width < 1 || ds->channels < 3)
    return fail(TNF_ERR_INVALID_ARGUMENT, "dataset: num_images=%d height=%d width=%d channels=%d", ds->num_images,
                ds->height, ds->width, ds->channels);
  if (!ds->images || !ds->camera_to_worlds || !ds->intrinsics)
    return fail(TNF_ERR_INVALID_ARGUMENT, "dataset: images/camera_to_worlds/intrinsics is null");
  if (num_rays == 0) return TNF_OK;
  if (!rand || !origins || !directions || !camera_indices)
    return fail(TNF_ERR_INVALID_ARGUMENT, "rand/origins/directions/camera_indices is null");
  tnf::SampleArgs a;
  a.rand = rand; a.images = ds->images; a.thermal = ds->thermal; a.c2w = ds->camera_to_worlds;
  a.intrinsics = ds->intrinsics; a.R = num_rays; a.N = ds->num_images; a.H = ds->height; a.W = ds->width;
  a.C = ds->channels; a.images_u8 = ds->images_uint8; a.thermal_u8 = ds->thermal_uint8;
  a.origins = origins; a.directions = directions; a.camera_indices = reinterpret_cast<long long*>(camera_indices);
  a.indices = reinterpret_cast<long long*>(indices); a.gt_rgb = gt_rgb; a.gt_thermal = gt_thermal;
  const int tb = 256;
  tnf::tnf_sample_batch_kernel<<<(unsigned)((num_rays + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "sample_batch launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}
