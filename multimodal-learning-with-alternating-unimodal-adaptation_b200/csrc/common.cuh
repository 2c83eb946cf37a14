// Shared helpers for libmla_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <atomic>
#include "../../include/mla_b200.h"

namespace mla {

struct DeviceInfo {
  int ok;          // 0 = not probed, 1 = usable, <0 = error code
  int device;
  int sm_count;
  int coop;
  int smem_optin;  // max opt-in dynamic shared memory per block
  int cc_major;
};

// Cached attributes of the current device (the library is used one-device-per-process).
const DeviceInfo& device_info();

extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Halo-resident 3x3 / stride-1 convolution path with resident weights (conv_strip16.cu: 64 -> 64 channels, fp16 operands).
// tiles == 0: not applicable, use the im2col kernel.
struct StripPlan {
  int Wp, TR, tiles_per_img, tiles;
};
StripPlan strip16_plan(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad);
int conv_strip16_run(int mode, const void* src16, const void* w16, float* out, int N, int H, int W, int accumulate,
                     float* stat_part, const float* out_scale, const StripPlan& pl, void* stream);

// tcgen05 attention forward for head width 64 (attention_tc.cu); `applicable` also honours MLA_ATTN_TC=0
bool attn_tc_applicable(int S, int Dh);
int attn_fwd_tc(const void* qkv16, const float* mask, float* out, float* stats, int B, int S, int H, float scale, void* stream);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Fixed-order block reduction (deterministic): warp shuffles, then warp 0 over the
// per-warp partials. `scratch` needs 32 floats. Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();            // scratch may still be read from a previous call
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? scratch[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

}  // namespace mla

#define MLA_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)
