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

}  // namespace

extern "C" size_t mla_head_ce_workspace_bytes(int B, int D, int C) {
  if (B < 1 || D < 4 || C < 1) return 0;
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
