// Test-time fusion + accuracy counters in ONE cooperative launch — replaces
// calculate_entropy / calculate_gating_weights[3] (reference main.py:65-106), the weighted
// sum (main.py:640-651) and the per-sample numpy argmax/counter loop (main.py:653-676),
// which costs ~8 device->host syncs per SAMPLE in the reference.
//
// The reference's "per-sample uncertainty" is in fact batch-global: softmax over dim=0 (the
// batch axis), summed over everything -> one scalar entropy per modality. That
// is what is computed here:
//   phase 1  each CTA loads its row chunk of all M logit matrices into smem (the only HBM
//            read), per-column partial (max, sum exp)
//   phase 2  combine partials -> column (max, S); entropy partial of own rows, literally
//            p*log(p) so that 0*log(0) = NaN propagates exactly as in the reference
//   phase 3  H_m, w = softmax(-H) with python max() NaN semantics; fused logits, argmax of
//            fused and of every modality, per-class counters
#include <math.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxM = 4;

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

__global__ void __launch_bounds__(kThreads, 1) fuse_eval_kernel(FuseParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  const int M = p.M, C = p.C, MC = p.M * p.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int nr = max(0, min(p.rows_per_cta, p.B - r0));
  float* s_x = smem;                                   // [M][rows_per_cta][C]
  float* s_gm = s_x + (size_t)M * p.rows_per_cta * C;  // [MC]
  float* s_gs = s_gm + MC;                             // [MC]
  float* s_red = s_gs + MC;                            // [32]
  __shared__ float s_w[kMaxM];

  const size_t mstride = (size_t)p.rows_per_cta * C;
  for (int m = 0; m < M; ++m) {
    const float* src = p.logits[m] + (size_t)r0 * C;
    for (int i = tid; i < nr * C; i += kThreads) s_x[m * mstride + i] = src[i];
  }
  __syncthreads();

  if (p.dynamic) {
    // phase 1: partial column stats over own rows
    for (int col = tid; col < MC; col += kThreads) {
      const int m = col / C, c = col - m * C;
      const float* x = s_x + m * mstride + c;
      float mx = -INFINITY;
      for (int r = 0; r < nr; ++r) mx = fmaxf(mx, x[(size_t)r * C]);
      float s = 0.f;
      for (int r = 0; r < nr; ++r) s += expf(x[(size_t)r * C] - mx);
      p.part_max[(size_t)blockIdx.x * MC + col] = mx;
      p.part_sum[(size_t)blockIdx.x * MC + col] = s;
    }
    grid.sync();
    // phase 2: global column stats (fixed order over CTAs), entropy partial of own rows
    for (int col = tid; col < MC; col += kThreads) {
      float gm = -INFINITY;
      for (int g = 0; g < (int)gridDim.x; ++g) gm = fmaxf(gm, p.part_max[(size_t)g * MC + col]);
      float gs = 0.f;
      for (int g = 0; g < (int)gridDim.x; ++g) {
        const float pm = p.part_max[(size_t)g * MC + col];
        const float ps = p.part_sum[(size_t)g * MC + col];
        if (ps > 0.f) gs += ps * expf(pm - gm);
      }
      s_gm[col] = gm;
      s_gs[col] = gs;
    }
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      float e = 0.f;
      for (int i = tid; i < nr * C; i += kThreads) {
        const int c = i % C;
        const float pr = expf(s_x[m * mstride + i] - s_gm[m * C + c]) / s_gs[m * C + c];
        e += pr * logf(pr);  // 0 * -inf = NaN, as in the reference
      }
      e = mla::block_sum(e, s_red);
      if (tid == 0) p.ent_part[(size_t)blockIdx.x * M + m] = (double)e;
    }
    grid.sync();
    // phase 3a: weights, computed identically by every CTA
    if (tid == 0) {
      float H[kMaxM];
      for (int m = 0; m < M; ++m) {
        double s = 0.0;
        for (int g = 0; g < (int)gridDim.x; ++g) s += p.ent_part[(size_t)g * M + m];
        H[m] = (float)(-s);
      }
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
  for (int r = warp; r < nr; r += kWarps) {
    const int b = r0 + r;
    const int lab = p.label ? (int)p.label[b] : -1;
    if (p.num && lane == 0 && lab >= 0 && lab < C) atomicAdd(p.num + lab, 1ull);
    for (int src = 0; src <= M; ++src) {
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
        if (p.hits && bi == lab) atomicAdd(p.hits + (size_t)src * C + lab, 1ull);
      }
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
  int grid = (B + 31) / 32;
  grid = grid < 1 ? 1 : (grid > di.sm_count ? di.sm_count : grid);
  int rpc = (B + grid - 1) / grid;
  grid = (B + rpc - 1) / rpc;
  size_t smem = ((size_t)M * rpc * C + 2 * (size_t)M * C + 32) * sizeof(float);
  if (smem > (size_t)di.smem_optin) return MLA_E_SHAPE;
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
    MLA_CUDA_TRY(cudaFuncSetAttribute(fuse_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin));
    s_smem_set.store((size_t)di.smem_optin, std::memory_order_relaxed);
  }
  void* args[] = {&prm};
  MLA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)fuse_eval_kernel, dim3(pl.grid), dim3(kThreads), args,
                                           pl.smem, static_cast<cudaStream_t>(stream)));
  mla::count_launch();
  return 0;
}
