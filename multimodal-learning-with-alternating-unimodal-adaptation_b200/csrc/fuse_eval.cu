// Test-time fusion + accuracy counters in ONE cooperative launch — replaces
// calculate_entropy / calculate_gating_weights[3] (reference main.py:65-106), the weighted
// sum (main.py:640-651) and the per-sample numpy argmax/counter loop (main.py:653-676),
// which costs ~8 device->host syncs per SAMPLE in the reference.
//
// The reference's "per-sample uncertainty" is in fact batch-global: softmax over dim=0 (the
// batch axis), summed over everything -> one scalar entropy per modality. That
// is what is computed here:
//   phase 1  each CTA loads its row chunk of all M logit matrices into smem (the only HBM
//            read), per-column partial (max, sum exp): every thread owns (column, row group),
//            the row-group partials are merged through shared memory in group order
//   phase 2  combine the CTA partials -> column (max, S): a warp per column, lane = CTA, fixed
//            shuffle tree; entropy partial of own rows, literally p*log(p) so that
//            0*log(0) = NaN propagates exactly as in the reference
//   phase 3  H_m (warp m, lane = CTA), w = softmax(-H) with python max() NaN semantics; fused
//            logits, argmax of fused and of every modality, per-class counters
// The grid is small on purpose (<= 64 CTAs of 1024 threads whenever the rows fit in shared memory):
// the op moves a few MB at most, so its time is launch latency + two grid barriers + the partial
// combines, all of which grow with the CTA count.
#include <math.h>
#include <algorithm>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxM = 4;
constexpr int kStaticReserve = 2048;   // the opt-in limit covers static + dynamic shared memory (the kernel has ~1 KB static)

struct FuseParams {
  const float* logits[kMaxM];
  float fixed_w[kMaxM];
  int M, B, C, dynamic, rows_per_cta;
  const int64_t* label;
  float* fused;
  float* w_out;
  float* h_out;
  int32_t* argmax;
  unsigned long long* hits;
  unsigned long long* num;
  float* part_max;   // [grid][M*C]
  float* part_sum;   // [grid][M*C]
  double* ent_part;  // [grid][M]
};

// merge of two (max, sum exp) partials; an empty partial is (-inf, 0)
__device__ __forceinline__ void ms_merge(float& m, float& s, float om, float os) {
  const float nm = fmaxf(m, om);
  const float a = (s > 0.f) ? s * expf(m - nm) : 0.f;
  const float b = (os > 0.f) ? os * expf(om - nm) : 0.f;
  m = nm;
  s = a + b;
}

__global__ void __launch_bounds__(kThreads, 1) fuse_eval_kernel(FuseParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  const int M = p.M, C = p.C, MC = p.M * p.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int nr = max(0, min(p.rows_per_cta, p.B - r0));
  const int G = (int)gridDim.x;
  float* s_x = smem;                                   // [M][rows_per_cta][C]
  float* s_gm = s_x + (size_t)M * p.rows_per_cta * C;  // [MC]
  float* s_gs = s_gm + MC;                             // [MC]
  float* s_pm = s_gs + MC;                             // [nrg][MC] row-group partial maxima
  float* s_ps = s_pm + kThreads;                       // [nrg][MC]              sums   (nrg * MC <= kThreads)
  float* s_red = s_ps + kThreads;                      // [32]
  int* s_lab = reinterpret_cast<int*>(s_red + 32);     // [rows_per_cta] labels (-1: none / out of range)
  __shared__ float s_w[kMaxM];
  __shared__ double s_ent[kMaxM][kWarps];

  const size_t mstride = (size_t)p.rows_per_cta * C;
  {
    const int n = nr * C;
    for (int m = 0; m < M; ++m) {
      const float* src = p.logits[m] + (size_t)r0 * C;
      float* dst = s_x + m * mstride;
      const bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && ((m * mstride) & 3) == 0;
      const int nv = vec ? (n >> 2) : 0;
      for (int i = tid; i < nv; i += kThreads)
        reinterpret_cast<float4*>(dst)[i] = __ldcs(reinterpret_cast<const float4*>(src) + i);
      for (int i = 4 * nv + tid; i < n; i += kThreads) dst[i] = __ldcs(src + i);
    }
  }
  for (int r = tid; r < nr; r += kThreads) {
    const long long l = p.label ? (long long)p.label[r0 + r] : -1;
    s_lab[r] = (l >= 0 && l < C) ? (int)l : -1;
  }
  __syncthreads();

  if (p.dynamic) {
    // phase 1: column stats over own rows. Thread = (column, row group rg of nrg): rows rg, rg + nrg, ...
    const int nrg = max(1, min(kThreads / MC, 32));
    for (int col0 = 0; col0 < MC; col0 += kThreads) {             // one pass unless M * C > 1024
      const int col = col0 + (nrg > 1 ? tid % MC : tid);
      const int rg = nrg > 1 ? tid / MC : 0;
      if (col < MC && rg < nrg) {
        const int m = col / C, c = col - m * C;
        const float* x = s_x + m * mstride + c;
        float mx = -INFINITY;
        for (int r = rg; r < nr; r += nrg) mx = fmaxf(mx, x[(size_t)r * C]);
        float s = 0.f;
        for (int r = rg; r < nr; r += nrg) s += expf(x[(size_t)r * C] - mx);
        if (nrg > 1) { s_pm[rg * MC + col] = mx; s_ps[rg * MC + col] = s; }
        else {
          p.part_max[(size_t)blockIdx.x * MC + col] = mx;
          p.part_sum[(size_t)blockIdx.x * MC + col] = s;
        }
      }
    }
    if (nrg > 1) {
      __syncthreads();
      if (tid < MC) {
        float mx = s_pm[tid], s = s_ps[tid];
        for (int rg = 1; rg < nrg; ++rg) ms_merge(mx, s, s_pm[rg * MC + tid], s_ps[rg * MC + tid]);
        p.part_max[(size_t)blockIdx.x * MC + tid] = mx;
        p.part_sum[(size_t)blockIdx.x * MC + tid] = s;
      }
    }
    grid.sync();
    // phase 2: global column stats. Warp per column, lane g holds CTAs g, g + 32, ... (merged in that order), then a fixed
    // shuffle tree over the lanes
    for (int col = warp; col < MC; col += kWarps) {
      float mx = -INFINITY, s = 0.f;
      for (int g = lane; g < G; g += 32)
        ms_merge(mx, s, __ldcg(p.part_max + (size_t)g * MC + col), __ldcg(p.part_sum + (size_t)g * MC + col));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const float os = __shfl_xor_sync(0xffffffffu, s, o);
        // both partners must form the same sum: order the operands by lane so that a + b is evaluated identically
        float lm = (lane & o) ? om : mx, ls = (lane & o) ? os : s;
        const float hm = (lane & o) ? mx : om, hs = (lane & o) ? s : os;
        ms_merge(lm, ls, hm, hs);
        mx = lm; s = ls;
      }
      if (lane == 0) { s_gm[col] = mx; s_gs[col] = s; }
    }
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      float e = 0.f;
      const float* x = s_x + m * mstride;
      // thread = (column c0 + k * cstep, rows r0 + k' * rstep): no division by C in the loop when C divides the block
      const int cstep = min(C, kThreads), rstep = max(1, kThreads / C);
      const int tc = tid % cstep, tr = tid / cstep;
      if (tr < rstep) {
        for (int c = tc; c < C; c += cstep) {
          const float gm = s_gm[m * C + c], gs = s_gs[m * C + c];
          for (int r = tr; r < nr; r += rstep) {
            const float pr = expf(x[(size_t)r * C + c] - gm) / gs;
            e += pr * logf(pr);  // 0 * -inf = NaN, as in the reference
          }
        }
      }
      e = mla::warp_sum(e);
      if (lane == 0) s_ent[m][warp] = (double)e;
    }
    __syncthreads();
    if (warp < M) {                                               // warp m: per-warp partials in warp order (fp64)
      double t = (lane < kWarps) ? s_ent[warp][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) p.ent_part[(size_t)blockIdx.x * M + warp] = t;
    }
    grid.sync();
    // phase 3a: weights, computed identically by every CTA. Warp m adds the CTA partials of modality m.
    if (warp < M) {
      double t = 0.0;
      for (int g = lane; g < G; g += 32) t += __ldcg(p.ent_part + (size_t)g * M + warp);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) s_red[warp] = (float)(-t);
    }
    __syncthreads();
    if (tid == 0) {
      float H[kMaxM];
      for (int m = 0; m < M; ++m) H[m] = s_red[m];
      float mx = H[0];  // python max(): keeps the first argument unless a later one compares greater
      for (int m = 1; m < M; ++m)
        if (H[m] > mx) mx = H[m];
      float g[kMaxM], sum = 0.f;
      for (int m = 0; m < M; ++m) { g[m] = expf(mx - H[m]); sum = __fadd_rn(sum, g[m]); }
      for (int m = 0; m < M; ++m) s_w[m] = __fdiv_rn(g[m], sum);
      if (blockIdx.x == 0 && p.h_out)
        for (int m = 0; m < M; ++m) p.h_out[m] = H[m];
    }
  } else {
    if (tid < M) s_w[tid] = p.fixed_w[tid];
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < M && p.w_out) p.w_out[tid] = s_w[tid];

  // phase 3b: fused = ((x0*w0 + x1*w1) + x2*w2) ... with the reference's roundings
  if (p.fused) {
    float* dst = p.fused + (size_t)r0 * C;
    for (int i = tid; i < nr * C; i += kThreads) {
      float v = __fmul_rn(s_x[i], s_w[0]);
      for (int m = 1; m < M; ++m) v = __fadd_rn(v, __fmul_rn(s_x[m * mstride + i], s_w[m]));
      dst[i] = v;
    }
  }
  // phase 3c: argmax (numpy semantics: first maximum; a NaN anywhere in the row -> the
  // softmax row is all-NaN -> index 0) and counters. Warp per (row, source).
  for (int u = warp; u < nr * (M + 1); u += kWarps) {
    const int r = u / (M + 1), src = u - r * (M + 1);
    const int b = r0 + r;
    const int lab = s_lab[r];
    if (src == 0 && p.num && lane == 0 && lab >= 0) atomicAdd(p.num + lab, 1ull);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    bool anynan = false;
    for (int c = lane; c < C; c += 32) {
      float v;
      if (src == 0) {
        v = __fmul_rn(s_x[(size_t)r * C + c], s_w[0]);
        for (int m = 1; m < M; ++m) v = __fadd_rn(v, __fmul_rn(s_x[m * mstride + (size_t)r * C + c], s_w[m]));
      } else {
        v = s_x[(src - 1) * mstride + (size_t)r * C + c];
      }
      if (v != v) anynan = true;
      if (v > best || (v == best && c < bi)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    anynan = __any_sync(0xffffffffu, anynan);
    if (anynan || bi == 0x7fffffff) bi = 0;
    if (lane == 0) {
      if (p.argmax) p.argmax[(size_t)src * p.B + b] = bi;
      if (p.hits && lab >= 0 && bi == lab) atomicAdd(p.hits + (size_t)src * C + lab, 1ull);
    }
  }
}

struct FusePlan {
  int grid, rows_per_cta;
  size_t smem, off_max, off_sum, off_ent, total;
};

int make_plan(int M, int B, int C, FusePlan* pl) {
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (M < 1 || M > kMaxM || B < 1 || C < 1) return MLA_E_SHAPE;
  // at most 32 CTAs (one lane per CTA in the partial combines) of >= 32 rows, unless the rows do not fit in shared memory
  const size_t fixed = (2 * (size_t)M * C + 2 * (size_t)kThreads + 32) * sizeof(float);   // + rows_per_cta labels, below
  if (fixed + (size_t)M * C * sizeof(float) > (size_t)di.smem_optin - kStaticReserve) return MLA_E_SHAPE;
  const int rpc_fit = (int)std::min<size_t>(((size_t)di.smem_optin - kStaticReserve - fixed) / (((size_t)M * C + 1) * sizeof(float)), (size_t)1 << 30);
  int rpc = std::max(32, (B + 63) / 64);                           // <= 64 CTAs: two partials per lane in the combines
  rpc = std::min(rpc, std::min(rpc_fit, B));
  int grid = (B + rpc - 1) / rpc;
  if (grid > di.sm_count) return MLA_E_SHAPE;                      // cooperative launch: one CTA per SM
  size_t smem = fixed + ((size_t)M * C + 1) * rpc * sizeof(float);
  pl->grid = grid; pl->rows_per_cta = rpc; pl->smem = smem;
  size_t off = 0;
  pl->off_max = off; off += mla::align_up((size_t)grid * M * C * 4, 256);
  pl->off_sum = off; off += mla::align_up((size_t)grid * M * C * 4, 256);
  pl->off_ent = off; off += mla::align_up((size_t)grid * M * 8, 256);
  pl->total = off;
  return 0;
}

}  // namespace

extern "C" size_t mla_fuse_eval_workspace_bytes(int M, int B, int C) {
  FusePlan pl;
  if (make_plan(M, B, C, &pl) != 0) return 0;
  return pl.total;
}

extern "C" int mla_fuse_eval(const float* const* logits, int M, int B, int C, int dynamic, const float* fixed_w,
                             const int64_t* label, float* fused, float* w_out, float* entropy_out,
                             int32_t* argmax, int64_t* hits, int64_t* num, void* ws, size_t ws_bytes, void* stream) {
  if (!logits) return MLA_E_BADARG;
  FusePlan pl;
  int rc = make_plan(M, B, C, &pl);
  if (rc != 0) return rc;
  if (!dynamic && !fixed_w) return MLA_E_BADARG;
  if ((hits || num) && !(hits && num && label)) return MLA_E_BADARG;
  if (ws == nullptr || ws_bytes < pl.total) return MLA_E_WORKSPACE;
  const mla::DeviceInfo& di = mla::device_info();
  if (!di.coop) return MLA_E_NODEVICE;
  FuseParams prm{};
  for (int m = 0; m < M; ++m) {
    if (!logits[m]) return MLA_E_BADARG;
    prm.logits[m] = logits[m];
    prm.fixed_w[m] = fixed_w ? fixed_w[m] : 0.f;
  }
  prm.M = M; prm.B = B; prm.C = C; prm.dynamic = dynamic ? 1 : 0; prm.rows_per_cta = pl.rows_per_cta;
  prm.label = label; prm.fused = fused; prm.w_out = w_out; prm.h_out = entropy_out; prm.argmax = argmax;
  prm.hits = reinterpret_cast<unsigned long long*>(hits);
  prm.num = reinterpret_cast<unsigned long long*>(num);
  char* w = static_cast<char*>(ws);
  prm.part_max = reinterpret_cast<float*>(w + pl.off_max);
  prm.part_sum = reinterpret_cast<float*>(w + pl.off_sum);
  prm.ent_part = reinterpret_cast<double*>(w + pl.off_ent);
  static std::atomic<size_t> s_smem_set{48 * 1024};
  if (pl.smem > s_smem_set.load(std::memory_order_relaxed)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(fuse_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin - kStaticReserve));
    s_smem_set.store((size_t)di.smem_optin, std::memory_order_relaxed);
  }
  void* args[] = {&prm};
  MLA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)fuse_eval_kernel, dim3(pl.grid), dim3(kThreads), args,
                                           pl.smem, static_cast<cudaStream_t>(stream)));
  mla::count_launch();
  return 0;
}
