// Fused multi-head attention core of the m3ae encoders (reference models/m3ae.py:103-121): softmax(scale * Q K^T with
// key-padding positions FILLED with -1e7) V, forward and backward, reading q/k/v in place from the qkv_linear output
// [B, S, 3, H, Dh] and writing the head-concatenated [B, S, H*Dh] layout the output projection consumes. The S x S score
// matrix never reaches HBM (the ATen path materialises it five times per layer: scores, scaled, filled, softmax, grad).
//
// Arithmetic: fp16 tensor-core MMAs (mma.sync.m16n8k16 fed by ldmatrix; operands rounded to fp16 when staged in shared
// memory — the same 10-bit mantissa as TF32 and as the tcgen05 fprop of the Linears — fp32 accumulate), fp32 online softmax.
// The probability / dS tiles go from accumulator to A-operand layout in registers. The backward pass is two deterministic
// kernels (dK/dV by key tile, dQ by query tile; no atomics) that recompute the probabilities from the saved row
// log-sum-exp. Gradients are far below fp16's normal range, so the backward scales dO by a power of two F chosen from
// max|dO| (found by the delta pre-pass) so that F * max|dO| lies in [128, 256), and divides dQ/dK/dV by F when storing:
// exact scaling, fp16 then behaves like TF32 over 2^-14..2^8 of the largest gradient element.
//
// Staging: q/k/v (and F * dO) are converted to fp16 ONCE per pass by a cast kernel; the attention kernels then stream
// 64-row tiles with cp.async into double-buffered shared memory, so the next tile's loads overlap the current tile's MMAs.
// The forward keeps (row max, row sum) instead of their log-sum-exp: with every key padded the scores are all -1e7 and
// -1e7 + log(S) is not representable in fp32, but exp(s - max) / sum is exact.
//
// These are warp-level mma.sync kernels, not tcgen05: attention is ~7 % of the encoder FLOPs; a TMEM-resident version is
// future work.
#include <cuda_fp16.h>
#include <math.h>
#include <algorithm>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kTile = 64;      // queries per CTA (16 per warp) and keys per inner tile
constexpr int kThreads = 128;

struct AttnParams {
  const __half* qkv;   // [B, S, 3, H, Dh] fp16
  const float* mask;   // [B, S] (> 0: padded key) or nullptr
  float* out;          // [B, S, H*Dh]
  float* lse;          // [B, H, S, 2]: (row max, row sum of exp)
  const __half* dout;  // [B, S, H*Dh] fp16 of F * dout
  const float* delta;  // [B, H, S] then one float: max|dout|
  float* dqkv;         // [B, S, 3, H, Dh]
  int B, S, H;
  float scale;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __half* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// A operand (16 x 16 at (r0, k0)) of a row-major fp16 tile
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __half* tile, int ld, int r0, int k0, int lane) {
  ldsm_x4(a, tile + (r0 + (lane & 15)) * ld + k0 + (lane >> 4) * 8);
}
// B operands of TWO adjacent n-tiles (n0, n0 + 8) for k0..k0+15 from a tile stored [n][k] (k contiguous): r[0..1], r[2..3]
__device__ __forceinline__ void load_b_nk(uint32_t (&r)[4], const __half* tile, int ld, int n0, int k0, int lane) {
  ldsm_x4(r, tile + (n0 + (lane & 7) + (lane >> 4) * 8) * ld + k0 + ((lane >> 3) & 1) * 8);
}
// the same from a tile stored [k][n] (n contiguous): transposing load
__device__ __forceinline__ void load_b_kn(uint32_t (&r)[4], const __half* tile, int ld, int n0, int k0, int lane) {
  ldsm_x4_t(r, tile + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + (lane >> 4) * 8);
}

// 64 rows x DH halfs from global (row r at src + (row0 + r) * row_stride, zero beyond `rows`) -> dst[r * (DH + 8) + c],
// asynchronously (cp.async, 16 bytes per request); the caller commits / waits
template <int DH>
__device__ __forceinline__ void load_tile(__half* dst, const __half* src, long long row_stride, int row0, int rows, int tid) {
  constexpr int V = DH / 8, LD = DH + 8;
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
  for (int i = tid; i < kTile * V; i += kThreads) {
    const int r = i / V, c = (i % V) * 8;
    const bool ok = row0 + r < rows;
    tc::cp_async16(d0 + (uint32_t)(r * LD + c) * 2u, src + (long long)(ok ? row0 + r : rows - 1) * row_stride + c, ok ? 16u : 0u);
  }
}

// dst16 = fp16(F * src), F = grad_scale(*amax) (1 when amax is null)
__device__ __forceinline__ float grad_scale(float amax);
__global__ void cast_scale16_kernel(const float* __restrict__ src, uint2* __restrict__ dst, long long n4,
                                    const float* __restrict__ amax) {
  const float F = amax != nullptr ? grad_scale(*amax) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
    dst[i] = make_uint2(pack_h2(v.x * F, v.y * F), pack_h2(v.z * F, v.w * F));
  }
}

// key state of a tile column: 0 = attend, 1 = padded (score filled with -1e7), 2 = beyond the sequence
__device__ __forceinline__ void load_key_state(float* ks, const float* mask, int b, int S, int key0, int tid) {
  if (tid < kTile) {
    const int key = key0 + tid;
    ks[tid] = key < S ? ((mask != nullptr && mask[(long long)b * S + key] > 0.f) ? 1.f : 0.f) : 2.f;
  }
}

// power of two F with F * amax in [128, 256)
__device__ __forceinline__ float grad_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);                       // amax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, max(-120, min(120, 8 - e)));
}

// --------------------------------------------------------------------------------------------------------------------
// forward: one CTA per (64-query tile, head, batch row); online softmax over 64-key tiles
// --------------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(AttnParams p) {
  constexpr int LD = DH + 8, NK = DH / 16, NT = DH / 8, TS = kTile * LD;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __half* Qs = reinterpret_cast<__half*>(sm_raw);
  __half* Ks = Qs + TS;                                  // 2 stages
  __half* Vs = Ks + 2 * TS;                              // 2 stages
  float* Kst = reinterpret_cast<float*>(Vs + 2 * TS);    // 2 stages
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = lane & 3, g = lane >> 2;
  const int S = p.S;
  const long long rs = 3LL * p.H * DH;
  const __half* base = p.qkv + (long long)b * S * rs + h * DH;
  load_tile<DH>(Qs, base, rs, qt * kTile, S, tid);
  load_tile<DH>(Ks, base + p.H * DH, rs, 0, S, tid);
  load_tile<DH>(Vs, base + 2 * p.H * DH, rs, 0, S, tid);
  load_key_state(Kst, p.mask, b, S, 0, tid);
  tc::cp_async_commit();
  uint32_t qa[NK][4];
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int nkt = (S + kTile - 1) / kTile;
  for (int kt = 0; kt < nkt; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nkt) {                      // prefetch the next key tile into the other stage (free since the last sync)
      load_tile<DH>(Ks + (cur ^ 1) * TS, base + p.H * DH, rs, (kt + 1) * kTile, S, tid);
      load_tile<DH>(Vs + (cur ^ 1) * TS, base + 2 * p.H * DH, rs, (kt + 1) * kTile, S, tid);
      load_key_state(Kst + (cur ^ 1) * kTile, p.mask, b, S, (kt + 1) * kTile, tid);
      tc::cp_async_commit();
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < NK; ++kk) load_a(qa[kk], Qs, LD, warp * 16, kk * 16, lane);
    }
    const __half* Kc = Ks + cur * TS;
    const __half* Vc = Vs + cur * TS;
    const float* Kstc = Kst + cur * kTile;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t kb[4];
        load_b_nk(kb, Kc, LD, jp * 16, kk * 16, lane);
        mma_f16(s[2 * jp], qa[kk], kb[0], kb[1]);
        mma_f16(s[2 * jp + 1], qa[kk], kb[2], kb[3]);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float st = Kstc[j * 8 + 2 * t + (e & 1)];
        float v = s[j][e] * p.scale;
        v = st == 0.f ? v : (st == 1.f ? -1e7f : -INFINITY);
        s[j][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float c0 = __expf(m0 - n0), c1 = __expf(m1 - n1);
    m0 = n0; m1 = n1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = __expf(s[j][0] - n0); s[j][1] = __expf(s[j][1] - n0);
      s[j][2] = __expf(s[j][2] - n1); s[j][3] = __expf(s[j][3] - n1);
      l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {        // 16 keys per step: the accumulator layout IS the A-operand layout
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < NK; ++jp) {
        uint32_t vb[4];
        load_b_kn(vb, Vc, LD, jp * 16, kk * 16, lane);
        mma_f16(o[2 * jp], pa, vb[0], vb[1]);
        mma_f16(o[2 * jp + 1], pa, vb[2], vb[3]);
      }
    }
    __syncthreads();                         // this stage is overwritten by the prefetch of the next iteration
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = qt * kTile + warp * 16 + g, r1 = r0 + 8;
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const long long os = (long long)p.H * DH;
  if (r0 < S) {
    float* orow = p.out + ((long long)b * S + r0) * os + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(orow + j * 8 + 2 * t) = make_float2(o[j][0] * i0, o[j][1] * i0);
    if (t == 0) *reinterpret_cast<float2*>(p.lse + 2 * (((long long)b * p.H + h) * S + r0)) = make_float2(m0, l0);
  }
  if (r1 < S) {
    float* orow = p.out + ((long long)b * S + r1) * os + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(orow + j * 8 + 2 * t) = make_float2(o[j][2] * i1, o[j][3] * i1);
    if (t == 0) *reinterpret_cast<float2*>(p.lse + 2 * (((long long)b * p.H + h) * S + r1)) = make_float2(m1, l1);
  }
}

// delta[b, h, s] = sum_d out * dout (one thread per (b, s, h)); *amax = max |dout| (caller zeroes it first)
template <int DH>
__global__ void attn_delta_kernel(const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ delta,
                                  float* __restrict__ amax, int B, int S, int H) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float mx = 0.f;
  if (i < (long long)B * S * H) {
    const int h = (int)(i % H);
    const long long bs = i / H;
    const int s = (int)(bs % S), b = (int)(bs / S);
    const float4* o = reinterpret_cast<const float4*>(out + i * DH);
    const float4* d = reinterpret_cast<const float4*>(dout + i * DH);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < DH / 4; ++k) {
      const float4 a = o[k], c = d[k];
      acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
    }
    delta[((long long)b * H + h) * S + s] = acc;
  }
  mx = mla::warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(mx));
}

// --------------------------------------------------------------------------------------------------------------------
// backward, dQ: one CTA per (64-query tile, head, batch row), loop over key tiles
// --------------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_kernel(AttnParams p) {
  constexpr int LD = DH + 8, NK = DH / 16, NT = DH / 8, TS = kTile * LD;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __half* Qs = reinterpret_cast<__half*>(sm_raw);
  __half* Os = Qs + TS;                                  // F * dO tile of these queries
  __half* Ks = Os + TS;                                  // 2 stages
  __half* Vs = Ks + 2 * TS;                              // 2 stages
  float* Kst = reinterpret_cast<float*>(Vs + 2 * TS);    // 2 stages
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = lane & 3, g = lane >> 2;
  const int S = p.S;
  const long long rs = 3LL * p.H * DH, os = (long long)p.H * DH;
  const __half* base = p.qkv + (long long)b * S * rs + h * DH;
  const float F = grad_scale(p.delta[(long long)p.B * p.H * S]);
  load_tile<DH>(Qs, base, rs, qt * kTile, S, tid);
  load_tile<DH>(Os, p.dout + (long long)b * S * os + h * DH, os, qt * kTile, S, tid);
  load_tile<DH>(Ks, base + p.H * DH, rs, 0, S, tid);
  load_tile<DH>(Vs, base + 2 * p.H * DH, rs, 0, S, tid);
  load_key_state(Kst, p.mask, b, S, 0, tid);
  tc::cp_async_commit();
  uint32_t qa[NK][4], da[NK][4];
  const int r0 = qt * kTile + warp * 16 + g, r1 = r0 + 8;
  const long long li = ((long long)b * p.H + h) * S;
  float M0 = 0.f, M1 = 0.f, I0 = 0.f, I1 = 0.f, D0 = 0.f, D1 = 0.f;
  if (r0 < S) { const float2 ml = *reinterpret_cast<const float2*>(p.lse + 2 * (li + r0)); M0 = ml.x; I0 = 1.f / ml.y; D0 = p.delta[li + r0] * F; }
  if (r1 < S) { const float2 ml = *reinterpret_cast<const float2*>(p.lse + 2 * (li + r1)); M1 = ml.x; I1 = 1.f / ml.y; D1 = p.delta[li + r1] * F; }
  float dq[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  const int nkt = (S + kTile - 1) / kTile;
  for (int kt = 0; kt < nkt; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nkt) {
      load_tile<DH>(Ks + (cur ^ 1) * TS, base + p.H * DH, rs, (kt + 1) * kTile, S, tid);
      load_tile<DH>(Vs + (cur ^ 1) * TS, base + 2 * p.H * DH, rs, (kt + 1) * kTile, S, tid);
      load_key_state(Kst + (cur ^ 1) * kTile, p.mask, b, S, (kt + 1) * kTile, tid);
      tc::cp_async_commit();
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < NK; ++kk) {
        load_a(qa[kk], Qs, LD, warp * 16, kk * 16, lane);
        load_a(da[kk], Os, LD, warp * 16, kk * 16, lane);
      }
    }
    const __half* Kc = Ks + cur * TS;
    const __half* Vc = Vs + cur * TS;
    const float* Kstc = Kst + cur * kTile;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t kb[4], vb[4];
        load_b_nk(kb, Kc, LD, jp * 16, kk * 16, lane);
        mma_f16(s[2 * jp], qa[kk], kb[0], kb[1]);
        mma_f16(s[2 * jp + 1], qa[kk], kb[2], kb[3]);
        load_b_nk(vb, Vc, LD, jp * 16, kk * 16, lane);
        mma_f16(dp[2 * jp], da[kk], vb[0], vb[1]);
        mma_f16(dp[2 * jp + 1], da[kk], vb[2], vb[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float st = Kstc[j * 8 + 2 * t + (e & 1)];
        const float M = e < 2 ? M0 : M1, I = e < 2 ? I0 : I1, D = e < 2 ? D0 : D1;
        // padded keys hold the CONSTANT -1e7 (torch.where): no gradient flows to their scores
        const float pr = st == 0.f ? __expf(s[j][e] * p.scale - M) * I : 0.f;
        s[j][e] = pr * (dp[j][e] - D) * p.scale;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < NK; ++jp) {
        uint32_t kb[4];
        load_b_kn(kb, Kc, LD, jp * 16, kk * 16, lane);
        mma_f16(dq[2 * jp], pa, kb[0], kb[1]);
        mma_f16(dq[2 * jp + 1], pa, kb[2], kb[3]);
      }
    }
    __syncthreads();
  }
  const float inv = 1.f / F;
  if (r0 < S) {
    float* row = p.dqkv + ((long long)b * S + r0) * rs + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(row + j * 8 + 2 * t) = make_float2(dq[j][0] * inv, dq[j][1] * inv);
  }
  if (r1 < S) {
    float* row = p.dqkv + ((long long)b * S + r1) * rs + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(row + j * 8 + 2 * t) = make_float2(dq[j][2] * inv, dq[j][3] * inv);
  }
}

// --------------------------------------------------------------------------------------------------------------------
// backward, dK and dV: one CTA per (64-key tile, head, batch row), loop over query tiles; everything transposed
// (rows = keys): S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
// --------------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kThreads, 3) attn_bwd_dkv_kernel(AttnParams p) {
  constexpr int LD = DH + 8, NK = DH / 16, NT = DH / 8, TS = kTile * LD;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __half* Ks = reinterpret_cast<__half*>(sm_raw);
  __half* Vs = Ks + TS;
  __half* Qs = Vs + TS;                                  // 2 stages
  __half* Os = Qs + 2 * TS;                              // 2 stages, F * dO
  float* Ms = reinterpret_cast<float*>(Os + 2 * TS);     // per stage: row max | 1 / row sum (0: no such query) | F * delta
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = lane & 3, g = lane >> 2;
  const int S = p.S;
  const long long rs = 3LL * p.H * DH, os = (long long)p.H * DH;
  const __half* base = p.qkv + (long long)b * S * rs + h * DH;
  const __half* obase = p.dout + (long long)b * S * os + h * DH;
  const long long li = ((long long)b * p.H + h) * S;
  const float F = grad_scale(p.delta[(long long)p.B * p.H * S]);
  auto load_rows = [&](float* dst, int q0) {
    if (tid < kTile) {
      const int q = q0 + tid;
      float2 ml = make_float2(0.f, 0.f);
      if (q < S) ml = *reinterpret_cast<const float2*>(p.lse + 2 * (li + q));
      dst[tid] = ml.x;
      dst[kTile + tid] = q < S ? 1.f / ml.y : 0.f;
      dst[2 * kTile + tid] = q < S ? p.delta[li + q] * F : 0.f;
    }
  };
  load_tile<DH>(Ks, base + p.H * DH, rs, kt * kTile, S, tid);
  load_tile<DH>(Vs, base + 2 * p.H * DH, rs, kt * kTile, S, tid);
  load_tile<DH>(Qs, base, rs, 0, S, tid);
  load_tile<DH>(Os, obase, os, 0, S, tid);
  load_rows(Ms, 0);
  tc::cp_async_commit();
  uint32_t ka[NK][4], va[NK][4];
  const int k0 = kt * kTile + warp * 16 + g, k1 = k0 + 8;      // this thread's key rows
  const float st0 = k0 < S ? ((p.mask != nullptr && p.mask[(long long)b * S + k0] > 0.f) ? 1.f : 0.f) : 2.f;
  const float st1 = k1 < S ? ((p.mask != nullptr && p.mask[(long long)b * S + k1] > 0.f) ? 1.f : 0.f) : 2.f;
  float dk[NT][4], dv[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const int nqt = (S + kTile - 1) / kTile;
  for (int qt = 0; qt < nqt; ++qt) {
    const int cur = qt & 1;
    if (qt + 1 < nqt) {
      load_tile<DH>(Qs + (cur ^ 1) * TS, base, rs, (qt + 1) * kTile, S, tid);
      load_tile<DH>(Os + (cur ^ 1) * TS, obase, os, (qt + 1) * kTile, S, tid);
      load_rows(Ms + (cur ^ 1) * 3 * kTile, (qt + 1) * kTile);
      tc::cp_async_commit();
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    __syncthreads();
    if (qt == 0) {
#pragma unroll
      for (int kk = 0; kk < NK; ++kk) {
        load_a(ka[kk], Ks, LD, warp * 16, kk * 16, lane);
        load_a(va[kk], Vs, LD, warp * 16, kk * 16, lane);
      }
    }
    const __half* Qc = Qs + cur * TS;
    const __half* Oc = Os + cur * TS;
    const float* Mc = Ms + cur * 3 * kTile;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t qb[4], ob[4];
        load_b_nk(qb, Qc, LD, jp * 16, kk * 16, lane);
        mma_f16(s[2 * jp], ka[kk], qb[0], qb[1]);
        mma_f16(s[2 * jp + 1], ka[kk], qb[2], qb[3]);
        load_b_nk(ob, Oc, LD, jp * 16, kk * 16, lane);
        mma_f16(dp[2 * jp], va[kk], ob[0], ob[1]);
        mma_f16(dp[2 * jp + 1], va[kk], ob[2], ob[3]);
      }
    }
    // P^T (rows = this warp's keys, columns = queries) and dS^T = P^T o (dP^T - delta) * scale (zero on padded keys)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = j * 8 + 2 * t + (e & 1);
        const float st = e < 2 ? st0 : st1;
        float pr = 0.f;
        if (st != 2.f) pr = __expf((st == 0.f ? s[j][e] * p.scale : -1e7f) - Mc[q]) * Mc[kTile + q];
        s[j][e] = pr;
        dp[j][e] = st == 0.f ? pr * (dp[j][e] - Mc[2 * kTile + q]) * p.scale : 0.f;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {        // 16 queries per step
      uint32_t pa[4], sa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      sa[0] = pack_h2(dp[2 * kk][0], dp[2 * kk][1]);
      sa[1] = pack_h2(dp[2 * kk][2], dp[2 * kk][3]);
      sa[2] = pack_h2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      sa[3] = pack_h2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < NK; ++jp) {
        uint32_t ob[4], qb[4];
        load_b_kn(ob, Oc, LD, jp * 16, kk * 16, lane);     // dV += P^T dO
        mma_f16(dv[2 * jp], pa, ob[0], ob[1]);
        mma_f16(dv[2 * jp + 1], pa, ob[2], ob[3]);
        load_b_kn(qb, Qc, LD, jp * 16, kk * 16, lane);     // dK += dS^T Q
        mma_f16(dk[2 * jp], sa, qb[0], qb[1]);
        mma_f16(dk[2 * jp + 1], sa, qb[2], qb[3]);
      }
    }
    __syncthreads();
  }
  const float inv = 1.f / F;
  if (k0 < S) {
    float* row = p.dqkv + ((long long)b * S + k0) * rs + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<float2*>(row + p.H * DH + j * 8 + 2 * t) = make_float2(dk[j][0] * inv, dk[j][1] * inv);
      *reinterpret_cast<float2*>(row + 2 * p.H * DH + j * 8 + 2 * t) = make_float2(dv[j][0] * inv, dv[j][1] * inv);
    }
  }
  if (k1 < S) {
    float* row = p.dqkv + ((long long)b * S + k1) * rs + h * DH;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      *reinterpret_cast<float2*>(row + p.H * DH + j * 8 + 2 * t) = make_float2(dk[j][2] * inv, dk[j][3] * inv);
      *reinterpret_cast<float2*>(row + 2 * p.H * DH + j * 8 + 2 * t) = make_float2(dv[j][2] * inv, dv[j][3] * inv);
    }
  }
}

template <int DH> constexpr size_t fwd_smem() { return (size_t)5 * kTile * (DH + 8) * 2 + 2 * kTile * 4; }
template <int DH> constexpr size_t dq_smem() { return (size_t)6 * kTile * (DH + 8) * 2 + 2 * kTile * 4; }
template <int DH> constexpr size_t dkv_smem() { return (size_t)6 * kTile * (DH + 8) * 2 + 6 * kTile * 4; }

bool attn_args_ok(int B, int S, int H, int Dh) {
  return B > 0 && S > 0 && H > 0 && (Dh == 32 || Dh == 64) && B <= 65535 && H <= 65535 &&
         (long long)B * S * H * Dh * 3 <= 0x3fffffffffLL;
}

unsigned cast_grid(long long n4) { return (unsigned)std::min<long long>((n4 + 255) / 256, 148LL * 16); }

template <int DH>
int run_fwd(const AttnParams& p, const float* qkv, cudaStream_t st) {
  static bool once = [] { return cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)fwd_smem<DH>()) == cudaSuccess; }();
  if (!once) return MLA_E_NODEVICE;
  const long long n4 = (long long)p.B * p.S * p.H * DH * 3 / 4;
  cast_scale16_kernel<<<cast_grid(n4), 256, 0, st>>>(qkv, reinterpret_cast<uint2*>(const_cast<__half*>(p.qkv)), n4, nullptr);
  MLA_CUDA_TRY(cudaGetLastError());
  if (mla::attn_tc_applicable(p.S, DH)) {        // head width 64, S <= 527: the tcgen05 / TMEM kernel (attention_tc.cu)
    const int rc = mla::attn_fwd_tc(p.qkv, p.mask, p.out, p.lse, p.B, p.S, p.H, p.scale, st);
    if (rc != 0) return rc;
    mla::count_launch(2);
    return 0;
  }
  dim3 grid((p.S + kTile - 1) / kTile, p.H, p.B);
  attn_fwd_kernel<DH><<<grid, kThreads, fwd_smem<DH>(), st>>>(p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch(2);
  return 0;
}

template <int DH>
int run_bwd(const AttnParams& p, const float* dout, float* delta, cudaStream_t st) {
  static bool once = [] {
    return cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dq_smem<DH>()) ==
               cudaSuccess &&
           cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkv_smem<DH>()) ==
               cudaSuccess;
  }();
  if (!once) return MLA_E_NODEVICE;
  const long long n = (long long)p.B * p.S * p.H;
  MLA_CUDA_TRY(cudaMemsetAsync(delta + n, 0, sizeof(float), st));
  attn_delta_kernel<DH><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p.out, dout, delta, delta + n, p.B, p.S, p.H);
  MLA_CUDA_TRY(cudaGetLastError());
  const long long n4 = n * DH / 4;
  cast_scale16_kernel<<<cast_grid(n4), 256, 0, st>>>(dout, reinterpret_cast<uint2*>(const_cast<__half*>(p.dout)), n4, delta + n);
  MLA_CUDA_TRY(cudaGetLastError());
  dim3 grid((p.S + kTile - 1) / kTile, p.H, p.B);
  attn_bwd_dkv_kernel<DH><<<grid, kThreads, dkv_smem<DH>(), st>>>(p);
  MLA_CUDA_TRY(cudaGetLastError());
  attn_bwd_dq_kernel<DH><<<grid, kThreads, dq_smem<DH>(), st>>>(p);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch(4);
  return 0;
}

size_t bwd_ws_floats(int B, int S, int H) { return mla::align_up((size_t)B * H * S + 4, 4); }

}  // namespace

extern "C" int mla_attention_forward(const float* qkv, const float* key_mask, float* out, float* stats, void* qkv16, int B,
                                     int S, int H, int Dh, float scale, void* stream) {
  if (!qkv || !out || !stats || !qkv16 || !mla::aligned16(qkv) || !mla::aligned16(out) || !mla::aligned16(qkv16) ||
      (reinterpret_cast<uintptr_t>(stats) & 7u))
    return MLA_E_BADARG;
  if (!attn_args_ok(B, S, H, Dh)) return MLA_E_SHAPE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  AttnParams p{};
  p.qkv = static_cast<const __half*>(qkv16); p.mask = key_mask; p.out = out; p.lse = stats;
  p.B = B; p.S = S; p.H = H; p.scale = scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return Dh == 64 ? run_fwd<64>(p, qkv, st) : run_fwd<32>(p, qkv, st);
}

extern "C" size_t mla_attention_backward_workspace_bytes(int B, int S, int H, int Dh) {
  if (!attn_args_ok(B, S, H, Dh)) return 0;
  return bwd_ws_floats(B, S, H) * sizeof(float) + (size_t)B * S * H * Dh * 2;
}

extern "C" int mla_attention_backward(const void* qkv16, const float* key_mask, const float* out, const float* dout,
                                      const float* stats, float* dqkv, int B, int S, int H, int Dh, float scale, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (!qkv16 || !out || !dout || !stats || !ws || !dqkv || !mla::aligned16(qkv16) || !mla::aligned16(out) ||
      !mla::aligned16(dout) || !mla::aligned16(dqkv) || !mla::aligned16(ws))
    return MLA_E_BADARG;
  if (!attn_args_ok(B, S, H, Dh)) return MLA_E_SHAPE;
  if (ws_bytes < mla_attention_backward_workspace_bytes(B, S, H, Dh)) return MLA_E_WORKSPACE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  float* delta = static_cast<float*>(ws);
  AttnParams p{};
  p.qkv = static_cast<const __half*>(qkv16); p.mask = key_mask; p.out = const_cast<float*>(out);
  p.lse = const_cast<float*>(stats); p.dout = reinterpret_cast<const __half*>(delta + bwd_ws_floats(B, S, H));
  p.delta = delta; p.dqkv = dqkv; p.B = B; p.S = S; p.H = H; p.scale = scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return Dh == 64 ? run_bwd<64>(p, dout, delta, st) : run_bwd<32>(p, dout, delta, st);
}
