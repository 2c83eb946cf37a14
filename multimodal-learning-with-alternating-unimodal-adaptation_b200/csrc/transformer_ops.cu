// Memory-bound pieces of the m3ae transformer block (reference models/m3ae.py:65-83,128-154): LayerNorm, exact GELU, the
// operand copies the tensor-core GEMMs consume and the bias gradients, fused so that every activation is read once and
// written once per pass:
//   * ln_fwd      x -> fp16 copy of LN(x) (GEMM operand) + TF32-rounded fp32 copy (weight-gradient operand) [+ plain fp32]
//   * ln_bwd      dx = resid + LN'(dy), per-block partial sums of dgamma / dbeta (fixed order -> deterministic)
//   * gelu_fwd    u -> fp16 + TF32-rounded gelu(u)
//   * gelu_bwd    du = dg * gelu'(u), TF32-rounded (it is the dy operand of the fc1 gradients) + column-sum partials (db)
//   * round_colsum dy -> TF32-rounded copy + column-sum partials (db of a Linear)
//   * cast_round  x -> fp16 + TF32-rounded copies
// All HBM-bound: coalesced float4 accesses, grids sized in multiples of the SM count.
#include <cuda_fp16.h>
#include <math.h>
#include <algorithm>
#include "common.cuh"

namespace {

__device__ __forceinline__ float tf32r(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float4 tf32r4(float4 v) { return make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w)); }

// ------------------------------------------------------------------------------------------------ LayerNorm forward
// one warp per row; NV = ceil(D / 128) float4 per lane
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, long long M, int D,
                                                     float* __restrict__ y, uint2* __restrict__ y16, float* __restrict__ y_r,
                                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps) {
    const float* xr = x + row * D;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = c < D ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mu = mla::warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rs = rsqrtf(mla::warp_sum(q) / D + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 o;
        o.x = (v[i].x - mu) * rs * g.x + b.x; o.y = (v[i].y - mu) * rs * g.y + b.y;
        o.z = (v[i].z - mu) * rs * g.z + b.z; o.w = (v[i].w - mu) * rs * g.w + b.w;
        if (y != nullptr) *reinterpret_cast<float4*>(y + row * D + c) = o;
        if (y16 != nullptr) y16[(row * D + c) >> 2] = make_uint2(h2(o.x, o.y), h2(o.z, o.w));
        if (y_r != nullptr) *reinterpret_cast<float4*>(y_r + row * D + c) = tf32r4(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// dx = resid + rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)); part[block][0][D] += dy * xhat,
// part[block][1][D] += dy. One warp per row, each block reduces its warps' partials through shared memory.
template <int NV>
__global__ void __launch_bounds__(256, (NV <= 6 ? 2 : 1)) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ resid,
                                                     long long M, int D, float* __restrict__ dx, float* __restrict__ part) {
  extern __shared__ float sh[];                 // [warps][2][D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long warps = (long long)gridDim.x * nw;
  for (long long row = (long long)blockIdx.x * nw + wib; row < M; row += warps) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], dg[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 xv = *reinterpret_cast<const float4*>(x + row * D + c);
        const float4 dv = *reinterpret_cast<const float4*>(dy + row * D + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
        ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
        dg[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
        s1 += dg[i].x + dg[i].y + dg[i].z + dg[i].w;
        s2 += dg[i].x * xh[i].x + dg[i].y * xh[i].y + dg[i].z * xh[i].z + dg[i].w * xh[i].w;
      } else {
        xh[i] = dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = mla::warp_sum(s1) / D;
    s2 = mla::warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        float4 o;
        o.x = rs * (dg[i].x - s1 - xh[i].x * s2); o.y = rs * (dg[i].y - s1 - xh[i].y * s2);
        o.z = rs * (dg[i].z - s1 - xh[i].z * s2); o.w = rs * (dg[i].w - s1 - xh[i].w * s2);
        if (resid != nullptr) {
          const float4 r = *reinterpret_cast<const float4*>(resid + row * D + c);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dx + row * D + c) = o;
      }
    }
  }
  // every warp parks its partials in shared memory; the block then adds them in warp order (deterministic)
  {
    float* mine = sh + (size_t)wib * 2 * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        *reinterpret_cast<float4*>(mine + c) = ag[i];
        *reinterpret_cast<float4*>(mine + D + c) = ab[i];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += sh[(size_t)w * 2 * D + i];
    part[(long long)blockIdx.x * 2 * D + i] = a;
  }
}

// out[c] = sum_b part[b * ld + c], c < W. Block = 32 columns x 32 row groups: group g adds rows g, g + 32, ... (4 loads in
// flight), then the groups are added in order through shared memory (deterministic).
__global__ void __launch_bounds__(1024) partial_reduce_kernel(const float* __restrict__ part, int nblk, int ld, int W,
                                                               float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < W) {
    int b = ty;
    for (; b + 96 < nblk; b += 128) {
      a0 += part[(long long)b * ld + c];
      a1 += part[(long long)(b + 32) * ld + c];
      a2 += part[(long long)(b + 64) * ld + c];
      a3 += part[(long long)(b + 96) * ld + c];
    }
    for (; b < nblk; b += 32) a0 += part[(long long)b * ld + c];
  }
  sm[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && c < W) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) a += sm[g][tx];
    out[c] = a;
  }
}

// ------------------------------------------------------------------------------------------------ elementwise + column sums
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_d(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * expf(-0.5f * u * u);
}

// KIND 0: cast_round (in -> out16, out_r); 1: gelu_fwd (u -> fp16 / tf32 gelu(u))
template <int KIND>
__global__ void ew_fwd_kernel(const float* __restrict__ in, uint2* __restrict__ out16, float* __restrict__ out_r, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(in + 4 * i);
    if (KIND == 1) v = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
    if (out16 != nullptr) out16[i] = make_uint2(h2(v.x, v.y), h2(v.z, v.w));
    if (out_r != nullptr) *reinterpret_cast<float4*>(out_r + 4 * i) = tf32r4(v);
  }
}

// KIND 0: out_r = tf32(dy); 1: out_r = tf32(dy * gelu'(u)). part[row chunk][N] = column sums of the UNROUNDED values.
// grid (column tiles of 4 * blockDim.x, row chunks); each thread owns 4 columns.
template <int KIND>
__global__ void ew_colsum_kernel(const float* __restrict__ dy, const float* __restrict__ u, float* __restrict__ out_r,
                                 float* __restrict__ part, long long M, int N, int rows_per_chunk) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (long long r = r0; r < r1; ++r) {
    float4 v = *reinterpret_cast<const float4*>(dy + r * N + c);
    if (KIND == 1) {
      const float4 uu = *reinterpret_cast<const float4*>(u + r * N + c);
      v.x *= gelu_d(uu.x); v.y *= gelu_d(uu.y); v.z *= gelu_d(uu.z); v.w *= gelu_d(uu.w);
    }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    *reinterpret_cast<float4*>(out_r + r * N + c) = tf32r4(v);
  }
  *reinterpret_cast<float4*>(part + (long long)blockIdx.y * N + c) = acc;
}

// power of two F with F * amax in [2^11, 2^12): fp16 then carries TF32's mantissa over 2^-26..1 of the largest element
__device__ __forceinline__ float operand_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);
  return ldexpf(1.f, max(-100, min(100, 12 - e)));
}

// pass 1 of the fp16 gradient operand: part[row chunk][N] = column sums of v = dy (KIND 0) or dy * gelu'(u) (KIND 1),
// *amax = max |v| (the caller zeroes it)
template <int KIND>
__global__ void colsum_amax_kernel(const float* __restrict__ dy, const float* __restrict__ u, float* __restrict__ part,
                                   float* __restrict__ amax, long long M, int N, int rows_per_chunk) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float mx = 0.f;
  if (c < N) {
    const long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
      float4 v = *reinterpret_cast<const float4*>(dy + r * N + c);
      if (KIND == 1) {
        const float4 uu = *reinterpret_cast<const float4*>(u + r * N + c);
        v.x *= gelu_d(uu.x); v.y *= gelu_d(uu.y); v.z *= gelu_d(uu.z); v.w *= gelu_d(uu.w);
      }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    *reinterpret_cast<float4*>(part + (long long)blockIdx.y * N + c) = acc;
  }
  mx = mla::warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(mx));
}

// pass 2: out16 = fp16(F * v), F = operand_scale(*amax); scale_io[1] = 1 / F for the GEMM epilogues
template <int KIND>
__global__ void cast_scaled16_kernel(const float* __restrict__ dy, const float* __restrict__ u, uint2* __restrict__ out16,
                                     float* __restrict__ scale_io, long long n4) {
  const float F = operand_scale(scale_io[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale_io[1] = 1.f / F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(dy + 4 * i);
    if (KIND == 1) {
      const float4 uu = *reinterpret_cast<const float4*>(u + 4 * i);
      v.x *= gelu_d(uu.x); v.y *= gelu_d(uu.y); v.z *= gelu_d(uu.z); v.w *= gelu_d(uu.w);
    }
    out16[i] = make_uint2(h2(v.x * F, v.y * F), h2(v.z * F, v.w * F));
  }
}

int ln_blocks(long long M, int per_sm = 2) {
  const mla::DeviceInfo& di = mla::device_info();
  return (int)std::min<long long>((M + 7) / 8, (long long)di.sm_count * per_sm);
}

int colsum_chunks(long long M, int N) {
  const mla::DeviceInfo& di = mla::device_info();
  const int col_tiles = (N + 1023) / 1024;
  const long long want = std::max(1, di.sm_count * 4 / col_tiles);
  return (int)std::max<long long>(1, std::min<long long>(want, (M + 15) / 16));
}

#define MLA_LAUNCH_OK()                       \
  do {                                        \
    MLA_CUDA_TRY(cudaGetLastError());         \
    mla::count_launch();                      \
  } while (0)

}  // namespace

extern "C" int mla_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps, long long M, int D,
                                     float* y, void* y16, float* y_r, float* mean, float* rstd, void* stream) {
  if (!x || !gamma || !beta || !mean || !rstd || M < 1 || D < 4 || (D & 3) || D > 1280 || !mla::aligned16(x) ||
      !mla::aligned16(gamma) || !mla::aligned16(beta) || !mla::aligned16(y) || !mla::aligned16(y_r) ||
      (reinterpret_cast<uintptr_t>(y16) & 7u))
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = ln_blocks(M, 8), nv = (D + 127) / 128;     // 64 registers: rows in flight, not occupancy, bound this pass
  uint2* h = static_cast<uint2*>(y16);
#define LN_FWD(NV) ln_fwd_kernel<NV><<<nb, 256, 0, st>>>(x, gamma, beta, eps, M, D, y, h, y_r, mean, rstd)
  if (nv <= 1) LN_FWD(1); else if (nv <= 3) LN_FWD(3); else if (nv <= 6) LN_FWD(6); else if (nv <= 8) LN_FWD(8); else LN_FWD(10);
#undef LN_FWD
  MLA_LAUNCH_OK();
  return 0;
}

extern "C" size_t mla_layernorm_backward_workspace_bytes(long long M, int D) {
  if (M < 1 || D < 4) return 0;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return 0;
  return (size_t)ln_blocks(M) * 2 * D * sizeof(float);
}

extern "C" int mla_layernorm_backward(const float* dy, const float* x, const float* mean, const float* rstd,
                                      const float* gamma, const float* resid, long long M, int D, float* dx, float* dgamma,
                                      float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || !ws || M < 1 || D < 4 || (D & 3) || D > 1280 ||
      !mla::aligned16(dy) || !mla::aligned16(x) || !mla::aligned16(gamma) || !mla::aligned16(resid) || !mla::aligned16(dx) ||
      !mla::aligned16(ws))
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (ws_bytes < mla_layernorm_backward_workspace_bytes(M, D)) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nb = ln_blocks(M), nv = (D + 127) / 128;
  float* part = static_cast<float*>(ws);
  const size_t sh = (size_t)8 * 2 * D * sizeof(float);
#define LN_BWD(NV)                                                                                                   \
  do {                                                                                                               \
    if (sh > 48 * 1024)                                                                                              \
      MLA_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));   \
    ln_bwd_kernel<NV><<<nb, 256, sh, st>>>(dy, x, mean, rstd, gamma, resid, M, D, dx, part);                         \
  } while (0)
  if (nv <= 1) LN_BWD(1); else if (nv <= 3) LN_BWD(3); else if (nv <= 6) LN_BWD(6); else if (nv <= 8) LN_BWD(8); else LN_BWD(10);
#undef LN_BWD
  MLA_LAUNCH_OK();
  // part[b] = [dgamma partial (D) | dbeta partial (D)]
  partial_reduce_kernel<<<(D + 31) / 32, 1024, 0, st>>>(part, nb, 2 * D, D, dgamma);
  MLA_LAUNCH_OK();
  partial_reduce_kernel<<<(D + 31) / 32, 1024, 0, st>>>(part + D, nb, 2 * D, D, dbeta);
  MLA_LAUNCH_OK();
  return 0;
}

static unsigned ew_grid(long long n4) {
  const mla::DeviceInfo& di = mla::device_info();
  return (unsigned)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, (long long)di.sm_count * 16));
}

extern "C" int mla_cast_round(const float* x, void* x16, float* x_r, long long n, int gelu, void* stream) {
  if (!x || (!x16 && !x_r) || n < 4 || (n & 3) || !mla::aligned16(x) || !mla::aligned16(x_r) ||
      (reinterpret_cast<uintptr_t>(x16) & 7u))
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gelu) ew_fwd_kernel<1><<<ew_grid(n / 4), 256, 0, st>>>(x, static_cast<uint2*>(x16), x_r, n / 4);
  else ew_fwd_kernel<0><<<ew_grid(n / 4), 256, 0, st>>>(x, static_cast<uint2*>(x16), x_r, n / 4);
  MLA_LAUNCH_OK();
  return 0;
}

extern "C" size_t mla_round_colsum_workspace_bytes(long long M, int N) {
  if (M < 1 || N < 4) return 0;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return 0;
  return (size_t)colsum_chunks(M, N) * N * sizeof(float);
}

extern "C" int mla_round_colsum(const float* dy, const float* u, float* out_r, float* colsum, long long M, int N, void* ws,
                                size_t ws_bytes, void* stream) {
  if (!dy || !out_r || !colsum || !ws || M < 1 || N < 4 || (N & 3) || !mla::aligned16(dy) || !mla::aligned16(u) ||
      !mla::aligned16(out_r) || !mla::aligned16(ws))
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (ws_bytes < mla_round_colsum_workspace_bytes(M, N)) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = colsum_chunks(M, N);
  const int rows = (int)((M + chunks - 1) / chunks);
  const int nchunks = (int)((M + rows - 1) / rows);
  float* part = static_cast<float*>(ws);
  dim3 grid((N + 1023) / 1024, nchunks);
  if (u != nullptr) ew_colsum_kernel<1><<<grid, 256, 0, st>>>(dy, u, out_r, part, M, N, rows);
  else ew_colsum_kernel<0><<<grid, 256, 0, st>>>(dy, nullptr, out_r, part, M, N, rows);
  MLA_LAUNCH_OK();
  partial_reduce_kernel<<<(N + 31) / 32, 1024, 0, st>>>(part, nchunks, N, N, colsum);
  MLA_LAUNCH_OK();
  return 0;
}

// The dy operand of a Linear's gradient GEMMs in fp16: out16 [M, N] = fp16(F * v), v = dy or dy * gelu'(u), F the power of
// two that puts max|v| in [2^11, 2^12); scale_io = 2 floats: [0] max|v| (scratch), [1] 1 / F (what mla_linear_dgrad16 /
// mla_linear_wgrad16 take as out_scale); colsum [N] = column sums of v (the bias gradient). Two passes over v.
extern "C" int mla_grad_operand16(const float* dy, const float* u, void* out16, float* colsum, float* scale_io, long long M,
                                  int N, void* ws, size_t ws_bytes, void* stream) {
  if (!dy || !out16 || !colsum || !scale_io || !ws || M < 1 || N < 4 || (N & 3) || !mla::aligned16(dy) || !mla::aligned16(u) ||
      (reinterpret_cast<uintptr_t>(out16) & 7u) || !mla::aligned16(ws))
    return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (ws_bytes < mla_round_colsum_workspace_bytes(M, N)) return MLA_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = colsum_chunks(M, N);
  const int rows = (int)((M + chunks - 1) / chunks);
  const int nchunks = (int)((M + rows - 1) / rows);
  float* part = static_cast<float*>(ws);
  MLA_CUDA_TRY(cudaMemsetAsync(scale_io, 0, sizeof(float), st));
  dim3 grid((N + 1023) / 1024, nchunks);
  if (u != nullptr) colsum_amax_kernel<1><<<grid, 256, 0, st>>>(dy, u, part, scale_io, M, N, rows);
  else colsum_amax_kernel<0><<<grid, 256, 0, st>>>(dy, nullptr, part, scale_io, M, N, rows);
  MLA_LAUNCH_OK();
  partial_reduce_kernel<<<(N + 31) / 32, 1024, 0, st>>>(part, nchunks, N, N, colsum);
  MLA_LAUNCH_OK();
  const long long n4 = M * N / 4;
  if (u != nullptr) cast_scaled16_kernel<1><<<ew_grid(n4), 256, 0, st>>>(dy, u, static_cast<uint2*>(out16), scale_io, n4);
  else cast_scaled16_kernel<0><<<ew_grid(n4), 256, 0, st>>>(dy, nullptr, static_cast<uint2*>(out16), scale_io, n4);
  MLA_LAUNCH_OK();
  return 0;
}
