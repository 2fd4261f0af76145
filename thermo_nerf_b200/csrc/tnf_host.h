// Host-side helpers shared by the translation units of libtnf_b200.so: the thread-local
// error string behind tnf_last_error() and model validation.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/tnf_b200.h"

namespace tnf {

inline thread_local char g_err[512] = "";

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// defined in tnf_forward.cu
int check_model(const TnfModel* m);

// defined in tnf_backward_tc.cu: the tensor-core field backward (one launch, weight gradients in tensor memory)
int launch_backward_field_tc(const TnfModel& m, const TnfRays& rays, const TnfSaved& sv, const TnfOutputGrads& go,
                             const TnfModelGrad& gr, cudaStream_t stream);

// cached per-thread device properties
inline int num_sms() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != dev_cached) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    dev_cached = dev;
  }
  return sms;
}

}  // namespace tnf
