// One modality turn of the shared head: Linear(D, C) + mean cross-entropy, forward and
// backward restricted to the head — reference main.py:432-435 with
// models/fusion_modules.py:19 and nn.CrossEntropyLoss (main.py:130).
//
//   kernel 1 (warp per sample): logits, log-softmax loss per row, dlogits
//   kernel 2 (thread per feature column): dW / db / feat_sum (reduction over the batch in a
//            fixed order) and dfeat (independent per sample), all in one launch
//
// feat_sum (sum_b feat) is emitted here because this kernel already streams feat; the GS
// projection consumes it (after the data-parallel all-reduce) instead of re-reading feat.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kBwdThreads = 128;
constexpr int kTile = 32;  // c-chunk (dW part) and b-chunk (dfeat part)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// smem: [kFwdWarps][D] feature rows, [kFwdWarps][Cpad] logits
__global__ void __launch_bounds__(kFwdThreads) head_fwd_kernel(
    const float* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ bias,
    const int64_t* __restrict__ label, int B, int D, int C, float grad_scale,
    float* __restrict__ logits, float* __restrict__ dlogits, float* __restrict__ rowloss) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cpad = (C + 31) & ~31;
  float* s_f = smem + (size_t)warp * D;
  float* s_l = smem + (size_t)kFwdWarps * D + (size_t)warp * Cpad;
  const int D4 = D >> 2;
  for (int b = blockIdx.x * kFwdWarps + warp; b < B; b += gridDim.x * kFwdWarps) {
    const float* fr = feat + (size_t)b * D;
    for (int j4 = lane; j4 < D4; j4 += 32) *reinterpret_cast<float4*>(s_f + 4 * j4) = ld4(fr + 4 * j4);
    __syncwarp();
    for (int c = 0; c < C; ++c) {
      const float* wr = W + (size_t)c * D;
      float acc = 0.f;
      for (int j4 = lane; j4 < D4; j4 += 32) {
        const float4 w = ld4(wr + 4 * j4);
        const float4 f = *reinterpret_cast<const float4*>(s_f + 4 * j4);
        acc = fmaf(w.x, f.x, acc); acc = fmaf(w.y, f.y, acc);
        acc = fmaf(w.z, f.z, acc); acc = fmaf(w.w, f.w, acc);
      }
      acc = mla::warp_sum(acc);
      if (lane == (c & 31)) s_l[c] = acc + (bias ? bias[c] : 0.f);
    }
    __syncwarp();
    // log-softmax over C
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, s_l[c]);
    m = mla::warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(s_l[c] - m);
    s = mla::warp_sum(s);
    const float lse = logf(s);
    // an out-of-range label (nn.CrossEntropyLoss raises a device assert for it) poisons this row's loss and gradient
    // with NaN instead of reading shared memory out of bounds: the error surfaces in the step's loss
    const long long lab64 = label[b];
    const bool lab_ok = lab64 >= 0 && lab64 < (long long)C;
    const int lab = lab_ok ? (int)lab64 : 0;
    for (int c = lane; c < C; c += 32) {
      const float l = s_l[c];
      if (logits) logits[(size_t)b * C + c] = l;
      if (dlogits) {
        const float pr = expf(l - m - lse);
        dlogits[(size_t)b * C + c] = lab_ok ? (pr - (c == lab ? 1.f : 0.f)) * grad_scale : __int_as_float(0x7fc00000);
      }
    }
    if (lane == 0) rowloss[b] = lab_ok ? -(s_l[lab] - m - lse) : __int_as_float(0x7fc00000);
    __syncwarp();
  }
}

// blockIdx.y <  ncc : dW rows [c0, c0+32), db, (c0 == 0: feat_sum, loss)
// blockIdx.y >= ncc : dfeat rows [b0, b0+32)
__global__ void __launch_bounds__(kBwdThreads) head_bwd_kernel(
    const float* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ dl,
    const float* __restrict__ rowloss, int B, int D, int C, int ncc, float* __restrict__ dW,
    float* __restrict__ db, float* __restrict__ dfeat, float* __restrict__ feat_sum,
    float* __restrict__ loss) {
  __shared__ float s_t[kTile][kTile + 1];
  __shared__ float s_red[32];
  const int j = blockIdx.x * kBwdThreads + threadIdx.x;
  const bool jok = j < D;
  if ((int)blockIdx.y < ncc) {
    const int c0 = blockIdx.y * kTile;
    const int ct = min(kTile, C - c0);
    float acc[kTile];
#pragma unroll
    for (int c = 0; c < kTile; ++c) acc[c] = 0.f;
    float fsum = 0.f;
    for (int b0 = 0; b0 < B; b0 += kTile) {
      const int bt = min(kTile, B - b0);
      __syncthreads();
      for (int i = threadIdx.x; i < kTile * kTile; i += kBwdThreads) {
        const int bb = i / kTile, cc = i % kTile;
        s_t[bb][cc] = (dl != nullptr && bb < bt && cc < ct) ? dl[(size_t)(b0 + bb) * C + c0 + cc] : 0.f;
      }
      __syncthreads();
      if (jok) {
        for (int bb = 0; bb < bt; ++bb) {
          const float f = feat[(size_t)(b0 + bb) * D + j];
          fsum += f;
#pragma unroll
          for (int c = 0; c < kTile; ++c) acc[c] = fmaf(s_t[bb][c], f, acc[c]);
        }
      }
    }
    if (jok) {
      if (dW) {
#pragma unroll
        for (int c = 0; c < kTile; ++c)
          if (c < ct) dW[(size_t)(c0 + c) * D + j] = acc[c];
      }
      if (c0 == 0 && feat_sum) feat_sum[j] = fsum;
    }
    if (blockIdx.x == 0) {
      if (db && dl && (int)threadIdx.x < ct) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dl[(size_t)b * C + c0 + threadIdx.x];
        db[c0 + threadIdx.x] = s;
      }
      if (c0 == 0 && loss) {
        float s = 0.f;
        for (int b = threadIdx.x; b < B; b += kBwdThreads) s += rowloss[b];
        s = mla::block_sum(s, s_red);
        if (threadIdx.x == 0) *loss = s / (float)B;
      }
    }
  } else {
    const int b0 = ((int)blockIdx.y - ncc) * kTile;
    const int bt = min(kTile, B - b0);
    float acc[kTile];
#pragma unroll
    for (int b = 0; b < kTile; ++b) acc[b] = 0.f;
    for (int c0 = 0; c0 < C; c0 += kTile) {
      const int ct = min(kTile, C - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < kTile * kTile; i += kBwdThreads) {
        const int bb = i / kTile, cc = i % kTile;
        s_t[cc][bb] = (bb < bt && cc < ct) ? dl[(size_t)(b0 + bb) * C + c0 + cc] : 0.f;
      }
      __syncthreads();
      if (jok) {
        for (int cc = 0; cc < ct; ++cc) {
          const float w = W[(size_t)(c0 + cc) * D + j];
#pragma unroll
          for (int b = 0; b < kTile; ++b) acc[b] = fmaf(s_t[cc][b], w, acc[b]);
        }
      }
    }
    if (jok) {
#pragma unroll
      for (int b = 0; b < kTile; ++b)
        if (b < bt) dfeat[(size_t)(b0 + b) * D + j] = acc[b];
    }
  }
}


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
  if (C <= kSC) return small_plan(B, D, C).bytes;
  return mla::align_up((size_t)B * C * 4, 256) + mla::align_up((size_t)B * 4, 256);
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
  float* dl = reinterpret_cast<float*>(ws);
  float* rowloss = reinterpret_cast<float*>(static_cast<char*>(ws) + mla::align_up((size_t)B * C * 4, 256));
  const bool need_bwd = dW || db || dfeat;

  const int Cpad = (C + 31) & ~31;
  const size_t smem1 = ((size_t)kFwdWarps * D + (size_t)kFwdWarps * Cpad) * sizeof(float);
  if (smem1 > (size_t)di.smem_optin) return MLA_E_SHAPE;
  static std::atomic<size_t> s_smem_set{48 * 1024};
  if (smem1 > s_smem_set.load(std::memory_order_relaxed)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
    s_smem_set.store((size_t)di.smem_optin, std::memory_order_relaxed);
  }
  int grid1 = (B + kFwdWarps - 1) / kFwdWarps;
  grid1 = grid1 > 4 * di.sm_count ? 4 * di.sm_count : grid1;
  head_fwd_kernel<<<grid1, kFwdThreads, smem1, st>>>(feat, W, bias, label, B, D, C, grad_scale, logits,
                                                    need_bwd ? dl : nullptr, rowloss);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();

  const int ncc = (C + kTile - 1) / kTile;
  const int nbc = dfeat ? (B + kTile - 1) / kTile : 0;
  // without backward outputs a single c-chunk still produces feat_sum / loss
  const int ny = (need_bwd ? ncc : 1) + nbc;
  dim3 grid2((D + kBwdThreads - 1) / kBwdThreads, ny);
  head_bwd_kernel<<<grid2, kBwdThreads, 0, st>>>(feat, W, need_bwd ? dl : nullptr, rowloss, B, D, C,
                                                need_bwd ? ncc : 1, dW, db, dfeat, feat_sum, loss);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
