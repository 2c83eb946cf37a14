// One modality turn of the shared head: Linear(D, C) + mean cross-entropy, forward and
// backward restricted to the head — reference main.py:432-435 with
// models/fusion_modules.py:19 and nn.CrossEntropyLoss (main.py:130).
//
//   C <= 16 (CREMA-D 6, IEMOCAP 4): memory-bound; head_rows / head_cols / head_reduce kernels (below)
//   C  > 16 (Food-101 101):         FMA-bound; three register-tiled fp32 GEMM launches + softmax / reductions (below)
// Every reduction has a fixed order (deterministic, identical on every data-parallel rank).
//
// feat_sum (sum_b feat) is emitted here because this kernel already streams feat; the GS
// projection consumes it (after the data-parallel all-reduce) instead of re-reading feat.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"

namespace {


__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------------------------------
// Small-C path (C <= 16: CREMA-D 6, IEMOCAP 4, MVSA 3): the head is memory-bound — feat is read once from HBM (logits) and
// once more from L2 (dW), dfeat is written once; W (C x D) lives in shared memory.
//   head_rows_kernel   a warp takes TWO samples at a time: logits for all C classes in registers (the W float4 from shared
//                      memory serves both samples), log-softmax / loss / dlogits, then dfeat = dlogits W streamed out
//   head_cols_kernel   grid (128-column chunks, batch splits): dW / feat_sum partial sums over the split's rows, rows dealt
//                      to the 8 warps, combined through shared memory in warp order; split 0 of chunk 0 also forms db / loss
//   head_reduce_kernel (only when the batch is split) adds the split partials in split order
// Every reduction has a fixed order: deterministic.
constexpr int kSC = 16;            // largest C of the small-C path
constexpr int kRowThreads = 256, kRowWarps = kRowThreads / 32;
constexpr int kColWarpsMax = 8;    // warps of head_cols_kernel (4 for the 16-class tile: its partial-sum staging must fit 48 KB)

template <int CT>   // CT = C rounded up to 4, 8 or 16 (register tile)
__global__ void __launch_bounds__(kRowThreads) head_rows_kernel(const float* __restrict__ feat, const float* __restrict__ W,
                                                                const float* __restrict__ bias,
                                                                const int64_t* __restrict__ label, int B, int D, int C,
                                                                float grad_scale, float* __restrict__ logits,
                                                                float* __restrict__ dlogits, float* __restrict__ rowloss,
                                                                float* __restrict__ dfeat, int w_in_smem) {
  extern __shared__ __align__(16) float s_w[];                   // [C][D] when w_in_smem
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D4 = D >> 2;
  if (w_in_smem) {
    for (int i = threadIdx.x; i < C * D4; i += kRowThreads) *reinterpret_cast<float4*>(s_w + 4 * i) = ld4(W + 4 * (size_t)i);
    __syncthreads();
  }
  const float* Wp = w_in_smem ? s_w : W;
  const int npairs = (B + 1) >> 1;
  for (int pr = blockIdx.x * kRowWarps + warp; pr < npairs; pr += gridDim.x * kRowWarps) {
    const int b0 = 2 * pr, b1 = min(2 * pr + 1, B - 1);
    const float* f0 = feat + (size_t)b0 * D;
    const float* f1 = feat + (size_t)b1 * D;
    float a0[CT], a1[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
    for (int j4 = lane; j4 < D4; j4 += 32) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(f0) + j4);
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(f1) + j4);
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        if (c < C) {
          const float4 w = *reinterpret_cast<const float4*>(Wp + (size_t)c * D + 4 * j4);
          a0[c] = fmaf(w.x, x0.x, a0[c]); a0[c] = fmaf(w.y, x0.y, a0[c]); a0[c] = fmaf(w.z, x0.z, a0[c]); a0[c] = fmaf(w.w, x0.w, a0[c]);
          a1[c] = fmaf(w.x, x1.x, a1[c]); a1[c] = fmaf(w.y, x1.y, a1[c]); a1[c] = fmaf(w.z, x1.z, a1[c]); a1[c] = fmaf(w.w, x1.w, a1[c]);
        }
      }
    }
    float d0[CT], d1[CT];                                         // logits, then dlogits (every lane holds all of them)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      if (c < C) {
        const float bc = bias ? __ldg(bias + c) : 0.f;
        d0[c] = mla::warp_sum(a0[c]) + bc;
        d1[c] = mla::warp_sum(a1[c]) + bc;
        m0 = fmaxf(m0, d0[c]); m1 = fmaxf(m1, d1[c]);
      } else {
        d0[c] = 0.f; d1[c] = 0.f;
      }
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c)
      if (c < C) { s0 += expf(d0[c] - m0); s1 += expf(d1[c] - m1); }
    const float lse0 = logf(s0), lse1 = logf(s1);
    const long long l0 = label[b0], l1 = label[b1];
    const bool ok0 = l0 >= 0 && l0 < C, ok1 = l1 >= 0 && l1 < C;
    const float nan = __int_as_float(0x7fc00000);
    float loss0 = nan, loss1 = nan;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      if (c < C) {
        const float lg0 = d0[c], lg1 = d1[c];
        if (logits != nullptr && lane == c) {
          logits[(size_t)b0 * C + c] = lg0;
          logits[(size_t)b1 * C + c] = lg1;
        }
        if (c == (int)l0 && ok0) loss0 = -(lg0 - m0 - lse0);
        if (c == (int)l1 && ok1) loss1 = -(lg1 - m1 - lse1);
        d0[c] = ok0 ? (expf(lg0 - m0 - lse0) - (c == (int)l0 ? 1.f : 0.f)) * grad_scale : nan;
        d1[c] = ok1 ? (expf(lg1 - m1 - lse1) - (c == (int)l1 ? 1.f : 0.f)) * grad_scale : nan;
        if (dlogits != nullptr && lane == c) {
          dlogits[(size_t)b0 * C + c] = d0[c];
          dlogits[(size_t)b1 * C + c] = d1[c];
        }
      }
    }
    if (lane == 0) { rowloss[b0] = loss0; rowloss[b1] = loss1; }
    if (dfeat != nullptr) {
      float4* o0 = reinterpret_cast<float4*>(dfeat + (size_t)b0 * D);
      float4* o1 = reinterpret_cast<float4*>(dfeat + (size_t)b1 * D);
      for (int j4 = lane; j4 < D4; j4 += 32) {
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          if (c < C) {
            const float4 w = *reinterpret_cast<const float4*>(Wp + (size_t)c * D + 4 * j4);
            r0.x = fmaf(d0[c], w.x, r0.x); r0.y = fmaf(d0[c], w.y, r0.y); r0.z = fmaf(d0[c], w.z, r0.z); r0.w = fmaf(d0[c], w.w, r0.w);
            r1.x = fmaf(d1[c], w.x, r1.x); r1.y = fmaf(d1[c], w.y, r1.y); r1.z = fmaf(d1[c], w.z, r1.z); r1.w = fmaf(d1[c], w.w, r1.w);
          }
        }
        __stcs(o0 + j4, r0);
        if (b1 != b0) __stcs(o1 + j4, r1);
      }
    }
  }
}

// grid (ceil(D / 128), S). Output: S == 1 -> dW / feat_sum directly; else partials part[s][(C + 1)][D] (row C = feat_sum).
template <int CT>
__global__ void __launch_bounds__(CT == 16 ? 128 : 256) head_cols_kernel(const float* __restrict__ feat, const float* __restrict__ dl,
                                                                const float* __restrict__ rowloss, int B, int D, int C,
                                                                int rows_per_split, float* __restrict__ dW,
                                                                float* __restrict__ db, float* __restrict__ feat_sum,
                                                                float* __restrict__ loss, float* __restrict__ part) {
  constexpr int kColWarps = CT == 16 ? 4 : 8, kColThreads = kColWarps * 32;
  __shared__ __align__(16) float s_acc[kColWarps][CT + 1][128];
  __shared__ float s_red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 128 + lane * 4;
  const int s = blockIdx.y, S = gridDim.y;
  const int r0 = s * rows_per_split, r1 = min(B, r0 + rows_per_split);
  float4 acc[CT], fs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < D) {
    int b = r0 + warp;
    for (; b + 3 * kColWarps < r1; b += 4 * kColWarps) {          // four rows in flight per warp
      float4 x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const float4*>(feat + (size_t)(b + q * kColWarps) * D + col));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        fs.x += x[q].x; fs.y += x[q].y; fs.z += x[q].z; fs.w += x[q].w;
        if (dl != nullptr) {
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            if (c < C) {
              const float g = __ldg(dl + (size_t)(b + q * kColWarps) * C + c);
              acc[c].x = fmaf(g, x[q].x, acc[c].x); acc[c].y = fmaf(g, x[q].y, acc[c].y);
              acc[c].z = fmaf(g, x[q].z, acc[c].z); acc[c].w = fmaf(g, x[q].w, acc[c].w);
            }
          }
        }
      }
    }
    for (; b < r1; b += kColWarps) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(feat + (size_t)b * D + col));
      fs.x += x.x; fs.y += x.y; fs.z += x.z; fs.w += x.w;
      if (dl != nullptr) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          if (c < C) {
            const float g = __ldg(dl + (size_t)b * C + c);
            acc[c].x = fmaf(g, x.x, acc[c].x); acc[c].y = fmaf(g, x.y, acc[c].y);
            acc[c].z = fmaf(g, x.z, acc[c].z); acc[c].w = fmaf(g, x.w, acc[c].w);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (c < C) *reinterpret_cast<float4*>(&s_acc[warp][c][lane * 4]) = acc[c];
  *reinterpret_cast<float4*>(&s_acc[warp][CT][lane * 4]) = fs;
  __syncthreads();
  // (C + 1) x 128 sums over the 8 warps, in warp order
  for (int i = threadIdx.x; i < (C + 1) * 128; i += kColThreads) {
    const int c = i / 128, j = i - c * 128;
    const int row = c < C ? c : CT;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kColWarps; ++w) t += s_acc[w][row][j];
    const int gj = blockIdx.x * 128 + j;
    if (gj < D) {
      if (S > 1) part[((size_t)s * (C + 1) + c) * D + gj] = t;
      else if (c < C) { if (dW) dW[(size_t)c * D + gj] = t; }
      else if (feat_sum) feat_sum[gj] = t;
    }
  }
  if (blockIdx.x == 0 && s == 0) {
    if (db != nullptr && dl != nullptr) {                          // db[c] = sum_b dl[b][c]: warp c, lanes over b, fixed tree
      for (int c = warp; c < C; c += kColWarps) {
        float t = 0.f;
        for (int b = lane; b < B; b += 32) t += dl[(size_t)b * C + c];
        t = mla::warp_sum(t);
        if (lane == 0) db[c] = t;
      }
    }
    if (loss != nullptr) {
      float t = 0.f;
      for (int b = threadIdx.x; b < B; b += kColThreads) t += rowloss[b];
      t = mla::block_sum(t, s_red);
      if (threadIdx.x == 0) *loss = t / (float)B;
    }
  }
}

__global__ void __launch_bounds__(256) head_reduce_kernel(const float* __restrict__ part, int S, int C, int D,
                                                          float* __restrict__ dW, float* __restrict__ feat_sum) {
  const long long n = (long long)(C + 1) * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = 0.f;
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldcg(part + (size_t)(s + q) * n + i);
#pragma unroll
      for (int q = 0; q < 8; ++q) t += v[q];
    }
    for (; s < S; ++s) t += __ldcg(part + (size_t)s * n + i);
    if (i < (long long)C * D) { if (dW) dW[i] = t; }
    else if (feat_sum) feat_sum[i - (long long)C * D] = t;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Large-C path (C > 16: Food-101's 101 classes): the head is three small fp32 GEMMs, bound by the FMA rate rather than HBM
// (6 B D C FLOP against 4 (2 B D + 3 C D + 2 B C) bytes: 150 FLOP / byte at C = 101). fp32 accuracy is part of the contract
// (rel 1e-5 against an fp64 evaluation), so they run on the CUDA cores: one register-tiled kernel (64 x 64 tile, 4 x 4 per
// thread, 16-deep k-steps, operands staged transposed in shared memory so that every thread reads two float4 per 16 FMAs)
// instantiated for the three operand layouts:
//   logits = feat W^T (+ b)          A = feat [B][D] (k contiguous), B = W [C][D] (k contiguous)
//   dfeat  = dlogits W               A = dlogits [B][C] (k contiguous), B = W [C][D] (n contiguous)
//   [dW; feat_sum] = [dlogits | 1]^T feat   A = dlogits^T (m contiguous) with a virtual all-ones row C, B = feat (n contiguous);
//                                    split over the batch, partials added in split order (deterministic)
// plus head_softmax_kernel (a warp per sample: loss, dlogits) and head_db_loss_kernel (a warp per class, fixed tree).
constexpr int kGT = 64, kGK = 16, kGPad = 4;

template <bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) head_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm,
                                                        int ldb, float* __restrict__ Cm, int ldc, int M, int N, int K,
                                                        int k_per_split, const float* __restrict__ bias_n, int ones_row) {
  __shared__ __align__(16) float As[kGK][kGT + kGPad];
  __shared__ __align__(16) float Bs[kGK][kGT + kGPad];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * kGT, n0 = blockIdx.x * kGT;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kb; k0 < ke; k0 += kGK) {
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {               // 64 x 16 elements of each operand, four per thread
      int m, k;
      if (A_KC) { k = t & 15; m = (t >> 4) + 16 * i; } else { m = t & 63; k = (t >> 6) + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < ke) v = (gm == ones_row) ? 1.f : (A_KC ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm]);
      ra[i] = v;
      int n, k2;
      if (B_NC) { n = t & 63; k2 = (t >> 6) + 4 * i; } else { k2 = t & 15; n = (t >> 4) + 16 * i; }
      const int gn = n0 + n, gk2 = k0 + k2;
      rb[i] = (gn < N && gk2 < ke) ? (B_NC ? Bm[(size_t)gk2 * ldb + gn] : Bm[(size_t)gn * ldb + gk2]) : 0.f;
    }
    __syncthreads();                            // the previous k-step has been consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (A_KC) As[t & 15][(t >> 4) + 16 * i] = ra[i]; else As[(t >> 6) + 4 * i][t & 63] = ra[i];
      if (B_NC) Bs[(t >> 6) + 4 * i][t & 63] = rb[i]; else Bs[t & 15][(t >> 4) + 16 * i] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* out = Cm + (size_t)blockIdx.z * M * ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) out[(size_t)gm * ldc + gn] = acc[i][j] + (bias_n != nullptr ? bias_n[gn] : 0.f);
    }
  }
}

// a warp per sample: (the split-K partial logits added in split order, + bias ->) logits row, log-softmax, loss, dlogits
__global__ void __launch_bounds__(256) head_softmax_kernel(float* __restrict__ logits, const float* __restrict__ lpart, int S,
                                                           const float* __restrict__ bias, const int64_t* __restrict__ label,
                                                           int B, int C, float grad_scale, float* __restrict__ dlogits,
                                                           float* __restrict__ rowloss) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float* row = logits + (size_t)b * C;
  if (lpart != nullptr) {
    const size_t n = (size_t)B * C;
    for (int c = lane; c < C; c += 32) {
      float t = 0.f;
      for (int sp = 0; sp < S; ++sp) t += __ldcg(lpart + (size_t)sp * n + (size_t)b * C + c);
      row[c] = t + (bias != nullptr ? bias[c] : 0.f);
    }
    __syncwarp();
  }
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
  m = mla::warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
  s = mla::warp_sum(s);
  const float lse = logf(s);
  const long long lab64 = label[b];
  const bool ok = lab64 >= 0 && lab64 < (long long)C;
  const int lab = ok ? (int)lab64 : 0;
  const float nan = __int_as_float(0x7fc00000);
  if (dlogits != nullptr)
    for (int c = lane; c < C; c += 32)
      dlogits[(size_t)b * C + c] = ok ? (expf(row[c] - m - lse) - (c == lab ? 1.f : 0.f)) * grad_scale : nan;
  if (lane == 0) rowloss[b] = ok ? -(row[lab] - m - lse) : nan;
}

// block c < C: db[c] = sum_b dl[b][c] (lanes over b, fixed tree over the 8 warps); block C: loss = mean of the row losses
__global__ void __launch_bounds__(256) head_db_loss_kernel(const float* __restrict__ dl, const float* __restrict__ rowloss, int B,
                                                           int C, float* __restrict__ db, float* __restrict__ loss) {
  __shared__ float s_red[32];
  const int c = blockIdx.x;
  float t = 0.f;
  if (c < C) {
    if (db == nullptr || dl == nullptr) return;
    for (int b = threadIdx.x; b < B; b += 256) t += dl[(size_t)b * C + c];
    t = mla::block_sum(t, s_red);
    if (threadIdx.x == 0) db[c] = t;
  } else {
    if (loss == nullptr) return;
    for (int b = threadIdx.x; b < B; b += 256) t += rowloss[b];
    t = mla::block_sum(t, s_red);
    if (threadIdx.x == 0) *loss = t / (float)B;
  }
}

// partial[s][rows][D] -> dW rows [0, C) and feat_sum (row C), splits added in order
__global__ void __launch_bounds__(256) head_split_reduce_kernel(const float* __restrict__ part, int S, int C, int D,
                                                                float* __restrict__ dW, float* __restrict__ feat_sum) {
  const long long n = (long long)(C + 1) * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < S; ++s) t += __ldcg(part + (size_t)s * n + i);
    if (i < (long long)C * D) { if (dW) dW[i] = t; }
    else if (feat_sum) feat_sum[i - (long long)C * D] = t;
  }
}

struct LargePlan { int S, k_per_split, S1, k1_per_split; size_t off_logits, off_rowloss, off_part, off_lpart, bytes; };
LargePlan large_plan(int B, int D, int C) {
  LargePlan pl;
  const mla::DeviceInfo& di = mla::device_info();
  const int sms = di.ok == 1 ? di.sm_count : 148;
  const int tiles = ((C + 1 + kGT - 1) / kGT) * ((D + kGT - 1) / kGT);
  int S = std::max(1, std::min((4 * sms + tiles - 1) / tiles, B / 128));      // >= 128 batch rows per split, ~4 CTAs per SM
  S = std::min(S, 32);
  pl.k_per_split = ((B + S - 1) / S + kGK - 1) / kGK * kGK;
  pl.S = (B + pl.k_per_split - 1) / pl.k_per_split;
  pl.off_logits = mla::align_up((size_t)B * C * 4, 256);                       // [0, off_logits): dlogits
  pl.off_rowloss = pl.off_logits + mla::align_up((size_t)B * C * 4, 256);
  pl.off_part = pl.off_rowloss + mla::align_up((size_t)B * 4, 256);
  // the logits GEMM is split over D when its (batch tile, class tile) grid alone would leave most SMs idle (small batches):
  // >= 64 columns per split; the partials are added in split order by the softmax kernel
  const int tiles1 = ((B + kGT - 1) / kGT) * ((C + kGT - 1) / kGT);
  int S1 = std::max(1, std::min(std::min((4 * sms + tiles1 - 1) / tiles1, D / 64), 32));
  pl.k1_per_split = ((D + S1 - 1) / S1 + kGK - 1) / kGK * kGK;
  pl.S1 = (D + pl.k1_per_split - 1) / pl.k1_per_split;
  pl.off_lpart = pl.off_part + mla::align_up((size_t)pl.S * (C + 1) * D * 4, 256);
  pl.bytes = pl.off_lpart + (pl.S1 > 1 ? mla::align_up((size_t)pl.S1 * B * C * 4, 256) : 0);
  return pl;
}

int run_large(const float* feat, const float* W, const float* bias, const int64_t* label, int B, int D, int C, float* logits,
              float* loss, float* dW, float* db, float* dfeat, float* feat_sum, float grad_scale, void* ws, cudaStream_t st) {
  const LargePlan pl = large_plan(B, D, C);
  char* base = static_cast<char*>(ws);
  float* dl = reinterpret_cast<float*>(base);
  float* lg = logits != nullptr ? logits : reinterpret_cast<float*>(base + pl.off_logits);
  float* rowloss = reinterpret_cast<float*>(base + pl.off_rowloss);
  float* part = reinterpret_cast<float*>(base + pl.off_part);
  const bool need_bwd = dW || db || dfeat;
  const dim3 blk(256);
  float* lpart = pl.S1 > 1 ? reinterpret_cast<float*>(base + pl.off_lpart) : nullptr;
  head_gemm_kernel<true, false><<<dim3((C + kGT - 1) / kGT, (B + kGT - 1) / kGT, pl.S1), blk, 0, st>>>(
      feat, D, W, D, lpart != nullptr ? lpart : lg, C, B, C, D, pl.k1_per_split, lpart != nullptr ? nullptr : bias, -1);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  head_softmax_kernel<<<(B + 7) / 8, 256, 0, st>>>(lg, lpart, pl.S1, bias, label, B, C, grad_scale, need_bwd ? dl : nullptr,
                                                   rowloss);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  if (dfeat != nullptr) {
    head_gemm_kernel<true, true><<<dim3((D + kGT - 1) / kGT, (B + kGT - 1) / kGT, 1), blk, 0, st>>>(
        dl, C, W, D, dfeat, D, B, D, C, C, nullptr, -1);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  if (dW != nullptr || feat_sum != nullptr) {
    // rows [0, C) = dlogits^T feat (only when gradients are wanted), last row = the virtual all-ones row = sum_b feat.
    // Always through the partial buffer: the reduction also splits the rows into dW and feat_sum.
    const int rows = dW != nullptr ? C + 1 : 1, ones = dW != nullptr ? C : 0;
    head_gemm_kernel<false, true><<<dim3((D + kGT - 1) / kGT, (rows + kGT - 1) / kGT, pl.S), blk, 0, st>>>(
        dl, C, feat, D, part, D, rows, D, B, pl.k_per_split, nullptr, ones);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
    const long long n = (long long)rows * D;
    head_split_reduce_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 4LL * mla::device_info().sm_count), 256, 0, st>>>(
        part, pl.S, rows - 1, D, dW, feat_sum);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  head_db_loss_kernel<<<C + 1, 256, 0, st>>>(need_bwd ? dl : nullptr, rowloss, B, C, db, loss);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

struct SmallPlan { int S, rows_per_split; size_t off_rowloss, off_part, bytes; };
SmallPlan small_plan(int B, int D, int C) {
  SmallPlan pl;
  const int chunks = (D + 127) / 128;
  const mla::DeviceInfo& di = mla::device_info();
  const int sms = di.ok == 1 ? di.sm_count : 148;
  int S = (4 * sms + chunks - 1) / chunks;                        // about four CTAs per SM
  S = std::max(1, std::min(S, B / (4 * kColWarpsMax)));           // >= 32 rows per split
  pl.rows_per_split = (B + S - 1) / S;
  pl.S = (B + pl.rows_per_split - 1) / pl.rows_per_split;
  pl.off_rowloss = mla::align_up((size_t)B * C * 4, 256);
  pl.off_part = pl.off_rowloss + mla::align_up((size_t)B * 4, 256);
  pl.bytes = pl.off_part + (pl.S > 1 ? mla::align_up((size_t)pl.S * (C + 1) * D * 4, 256) : 0);
  return pl;
}

template <int CT>
int run_small(const float* feat, const float* W, const float* bias, const int64_t* label, int B, int D, int C, float* logits,
              float* loss, float* dW, float* db, float* dfeat, float* feat_sum, float grad_scale, void* ws, cudaStream_t st) {
  const mla::DeviceInfo& di = mla::device_info();
  const SmallPlan pl = small_plan(B, D, C);
  float* dl = reinterpret_cast<float*>(ws);
  float* rowloss = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.off_rowloss);
  float* part = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.off_part);
  const bool need_bwd = dW || db || dfeat;
  const size_t wbytes = (size_t)C * D * 4;
  const int w_in_smem = wbytes <= (size_t)di.smem_optin - 1024 ? 1 : 0;
  static std::atomic<size_t> s_set{48 * 1024};
  if (w_in_smem && wbytes > s_set.load(std::memory_order_relaxed)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(head_rows_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.smem_optin - 1024));
    s_set.store((size_t)di.smem_optin, std::memory_order_relaxed);
  }
  const int npairs = (B + 1) / 2;
  int grid1 = (npairs + kRowWarps - 1) / kRowWarps;
  // staging W costs C * D * 4 bytes per CTA: few, fat CTAs when there is little work per CTA
  grid1 = std::max(1, std::min(grid1, (w_in_smem && wbytes > 64 * 1024 ? 1 : 2) * di.sm_count));
  head_rows_kernel<CT><<<grid1, kRowThreads, w_in_smem ? wbytes : 0, st>>>(feat, W, bias, label, B, D, C, grad_scale, logits,
                                                                          need_bwd ? dl : nullptr, rowloss, dfeat, w_in_smem);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  dim3 grid2((D + 127) / 128, pl.S);
  head_cols_kernel<CT><<<grid2, CT == 16 ? 128 : 256, 0, st>>>(feat, (dW || db) ? dl : nullptr, rowloss, B, D, C, pl.rows_per_split, dW, db,
                                                    feat_sum, loss, part);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  if (pl.S > 1 && (dW || feat_sum)) {
    const long long n = (long long)(C + 1) * D;
    head_reduce_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 2LL * di.sm_count), 256, 0, st>>>(part, pl.S, C, D, dW, feat_sum);
    MLA_CUDA_TRY(cudaGetLastError());
    mla::count_launch();
  }
  return 0;
}

}  // namespace

extern "C" size_t mla_head_ce_workspace_bytes(int B, int D, int C) {
  if (B < 1 || D < 4 || C < 1) return 0;
  // the larger of the two plans: MLA_HEAD_SMALL=0 may route a small C through the tiled path
  const size_t large = large_plan(B, D, C).bytes;
  return C <= kSC ? std::max(small_plan(B, D, C).bytes, large) : large;
}

extern "C" int mla_head_ce(const float* feat, const float* W, const float* bias, const int64_t* label, int B, int D,
                           int C, float* logits, float* loss, float* dW, float* db, float* dfeat, float* feat_sum,
                           float grad_scale, void* ws, size_t ws_bytes, void* stream) {
  if (!feat || !W || !label) return MLA_E_BADARG;
  if (B < 1 || D < 4 || (D & 3) || C < 1 || C > 1024) return MLA_E_SHAPE;
  if (!mla::aligned16(feat) || !mla::aligned16(W) || !mla::aligned16(ws)) return MLA_E_BADARG;
  const size_t need = mla_head_ce_workspace_bytes(B, D, C);
  if (ws == nullptr || ws_bytes < need) return MLA_E_WORKSPACE;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool small_off = [] { const char* e = getenv("MLA_HEAD_SMALL"); return e != nullptr && e[0] == '0'; }();
  if (C <= kSC && !small_off) {
    if (C <= 4) return run_small<4>(feat, W, bias, label, B, D, C, logits, loss, dW, db, dfeat, feat_sum, grad_scale, ws, st);
    if (C <= 8) return run_small<8>(feat, W, bias, label, B, D, C, logits, loss, dW, db, dfeat, feat_sum, grad_scale, ws, st);
    return run_small<16>(feat, W, bias, label, B, D, C, logits, loss, dW, db, dfeat, feat_sum, grad_scale, ws, st);
  }
  return run_large(feat, W, bias, label, B, D, C, logits, loss, dW, db, dfeat, feat_sum, grad_scale, ws, st);
}
