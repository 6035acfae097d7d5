// ctx.hpp -- context object behind the C ABI: device, stream, pre-allocated scratch, pinned staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "kernels.cuh"

#define RGBID_CUDA_TRY(expr)                                        \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return RGBID_ERR_CUDA_BASE + (int)_e;    \
  } while (0)

struct rgbid_ctx {
  int device;
  int num_sms;
  cudaStream_t stream;
  bool own_stream;
  long long launches;
  // device scratch for the drop-in (un-fused) entry points
  double* d_partials;          // [num_sms][32]
  unsigned int* d_counter;     // last-block ticket
  unsigned int* d_counts;      // visibility counters [4]
  double* d_out;               // 32 doubles
  rgbid::ScaleState* d_scale;
  // pinned host staging
  void* h_small;               // 4 KiB for scalar read-backs
  void* h_stage;               // grows on demand for host-image uploads
  size_t h_stage_bytes;
  void* d_stage;               // device side of host-image uploads
  size_t d_stage_bytes;

  rgbid::LaunchCtx L() { return rgbid::LaunchCtx{stream, &launches, num_sms}; }
};

namespace rgbid {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// pinned staging of at least `bytes`
int ctx_reserve_host_stage(rgbid_ctx* ctx, size_t bytes);
int ctx_reserve_device_stage(rgbid_ctx* ctx, size_t bytes);
int check_last(rgbid_ctx* ctx);

}  // namespace rgbid
