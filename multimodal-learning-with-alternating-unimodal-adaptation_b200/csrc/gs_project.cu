// Fused GSPlugin projection — replaces the body of GSPlugin.before_update
// (reference utils/utils.py:34-41). One cooperative launch:
//
//   phase 0  copy grad_w -> workspace (so the in-place projection can't race);
//            raw-feature path only: 2-stage deterministic batch-sum of feat -> r
//   phase A  k = P r^T          each CTA streams its row slice of P from HBM ONCE, keeps it
//                               in shared memory for the remaining phases
//   phase B  P' = P - (k k^T) ./ (alpha + k r)   in shared memory, + sum(P'^2) partials
//   phase C  P = P' / ||P'||_F  written to HBM once; grad_w = grad_w @ P^T from the
//                               smem-resident rows (register-blocked dot products)
//
// HBM traffic is the algorithmic minimum 4*(B*D + 2*D*D + 2*C*D) bytes: P is read once and
// written once. All reductions have a fixed order, so every rank of a data-parallel job
// that feeds identical (P, feat_sum, grad_w) computes bit-identical results.
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kGT = 8;        // grad rows staged per tile in phase C
constexpr int kChunk = 1024;  // columns held in registers per pass of phase C (8 float4/lane)

struct GsParams {
  float* P;
  const float* feat;
  const float* feat_sum;
  float inv_batch;
  float alpha;
  float* grad_w;
  int B, D, C, mode;
  int rows_per_cta;
  int nb;           // batch chunks of the raw-feature reduction
  int rows_per_nb;  // rows per batch chunk
  float* ws_r;      // [D]
  float* ws_k;      // [D]
  float* ws_part;   // [nb][D]
  double* ws_norm;  // [grid]
  float* ws_g;      // [C][D]
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__global__ void __launch_bounds__(kThreads, 1) gs_project_kernel(GsParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, D4 = p.D >> 2;
  float* s_r = smem;                      // [D]
  float* s_k = s_r + D;                   // [D]
  float* s_P = s_k + D;                   // [rows_per_cta][D]
  float* s_G = s_P + (size_t)p.rows_per_cta * D;  // [kGT][D]
  float* s_red = s_G + (size_t)kGT * D;   // [32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = blockIdx.x * kThreads + tid;
  const int gthreads = gridDim.x * kThreads;
  const int row0 = blockIdx.x * p.rows_per_cta;
  const int nrows = max(0, min(p.rows_per_cta, D - row0));

  // ---------------- phase 0
  if (p.grad_w != nullptr) {
    const int n4 = p.C * D4;
    for (int i = gtid; i < n4; i += gthreads) st4(p.ws_g + 4 * (size_t)i, ld4(p.grad_w + 4 * (size_t)i));
  }
  if (p.feat != nullptr) {
    // stage 1: unit = (128-column chunk, batch chunk); one warp per unit, float4 per lane.
    const int cchunks = (D + 127) / 128;
    const int units = cchunks * p.nb;
    const int gwarp = blockIdx.x * kWarps + warp, gwarps = gridDim.x * kWarps;
    for (int u = gwarp; u < units; u += gwarps) {
      const int cc = u % cchunks, bc = u / cchunks;
      const int col = cc * 128 + lane * 4;
      const int b0 = bc * p.rows_per_nb, b1 = min(p.B, b0 + p.rows_per_nb);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < D) {
        const float* src = p.feat + col;
        int b = b0;
        for (; b + 8 <= b1; b += 8) {
          float4 v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = ld4(src + (size_t)(b + q) * D);
#pragma unroll
          for (int q = 0; q < 8; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        }
        for (; b < b1; ++b) {
          float4 v = ld4(src + (size_t)b * D);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(p.ws_part + (size_t)bc * D + col, acc);
      }
    }
    grid.sync();
    // stage 2: r[j] = inv_batch * sum_bc part[bc][j]  (fixed order over bc)
    for (int j = gtid; j < D; j += gthreads) {
      float s = 0.f;
      for (int bc = 0; bc < p.nb; ++bc) s += p.ws_part[(size_t)bc * D + j];
      p.ws_r[j] = s * p.inv_batch;
    }
    grid.sync();
    for (int j = tid; j < D; j += kThreads) s_r[j] = p.ws_r[j];
  } else {
    for (int j = tid; j < D; j += kThreads) s_r[j] = p.feat_sum[j] * p.inv_batch;
  }
  __syncthreads();

  // ---------------- phase A: k_i = sum_j P_ij r_j ; P rows -> smem
  for (int lr = warp; lr < nrows; lr += kWarps) {
    const float* src = p.P + (size_t)(row0 + lr) * D;
    float* dst = s_P + (size_t)lr * D;
    float acc = 0.f;
    int j4 = lane;
    for (; j4 + 7 * 32 < D4; j4 += 8 * 32) {      // 8 independent 512-byte row segments in flight per warp
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ld4(src + 4 * (j4 + 32 * u));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 r = ld4(s_r + 4 * (j4 + 32 * u));
        st4(dst + 4 * (j4 + 32 * u), v[u]);
        acc = fmaf(v[u].x, r.x, acc); acc = fmaf(v[u].y, r.y, acc);
        acc = fmaf(v[u].z, r.z, acc); acc = fmaf(v[u].w, r.w, acc);
      }
    }
    for (; j4 < D4; j4 += 32) {
      float4 v = ld4(src + 4 * j4);
      float4 r = ld4(s_r + 4 * j4);
      st4(dst + 4 * j4, v);
      acc = fmaf(v.x, r.x, acc); acc = fmaf(v.y, r.y, acc);
      acc = fmaf(v.z, r.z, acc); acc = fmaf(v.w, r.w, acc);
    }
    acc = mla::warp_sum(acc);
    if (lane == 0) p.ws_k[row0 + lr] = acc;
  }
  grid.sync();

  // ---------------- phase B: elementwise update in smem + sum of squares
  for (int j = tid; j < D; j += kThreads) s_k[j] = p.ws_k[j];
  __syncthreads();
  float scal_den = 0.f;
  if (p.mode == 1) {  // canonical OWM: scalar denominator alpha + r.k (same order in every CTA)
    float part = 0.f;
    for (int j = tid; j < D; j += kThreads) part = fmaf(s_r[j], s_k[j], part);
    scal_den = __fadd_rn(p.alpha, mla::block_sum(part, s_red));
  }
  float sq = 0.f;
  {
    const int n4 = nrows * D4;
    for (int i = tid; i < n4; i += kThreads) {
      const int lr = i / D4, j4 = i - lr * D4;
      const float ki = s_k[row0 + lr];
      float4 pv = ld4(s_P + (size_t)lr * D + 4 * j4);
      const float4 kv = ld4(s_k + 4 * j4);
      const float4 rv = ld4(s_r + 4 * j4);
      // Same operation order and roundings as utils.py:36 (no FMA contraction):
      //   P - (k_i*k_j) / (alpha + k_i*r_j)
#define MLA_GS_UPD(c)                                                                       \
      {                                                                                     \
        const float den = (p.mode == 0) ? __fadd_rn(p.alpha, __fmul_rn(ki, rv.c)) : scal_den; \
        pv.c = __fsub_rn(pv.c, __fdiv_rn(__fmul_rn(ki, kv.c), den));                        \
        sq = fmaf(pv.c, pv.c, sq);                                                          \
      }
      MLA_GS_UPD(x) MLA_GS_UPD(y) MLA_GS_UPD(z) MLA_GS_UPD(w)
#undef MLA_GS_UPD
      st4(s_P + (size_t)lr * D + 4 * j4, pv);
    }
  }
  {
    const float bs = mla::block_sum(sq, s_red);
    if (tid == 0) p.ws_norm[blockIdx.x] = (double)bs;
  }
  grid.sync();

  // ---------------- phase C: normalise, write P, project the gradient
  double tot = 0.0;
  for (int c = 0; c < (int)gridDim.x; ++c) tot += p.ws_norm[c];  // fixed order, identical everywhere
  const float nrm = (float)sqrt(tot);
  {
    const int n4 = nrows * D4;
    for (int i = tid; i < n4; i += kThreads) {
      const int lr = i / D4, j4 = i - lr * D4;
      float4 pv = ld4(s_P + (size_t)lr * D + 4 * j4);
      pv.x = __fdiv_rn(pv.x, nrm); pv.y = __fdiv_rn(pv.y, nrm);
      pv.z = __fdiv_rn(pv.z, nrm); pv.w = __fdiv_rn(pv.w, nrm);
      st4(s_P + (size_t)lr * D + 4 * j4, pv);
      st4(p.P + (size_t)(row0 + lr) * D + 4 * j4, pv);
    }
  }
  if (p.grad_w == nullptr) return;
  __syncthreads();

  // grad_w[c][i] = sum_j G[c][j] * P[i][j].  Warp owns a pair of rows; per 1024-column
  // chunk the two P rows sit in registers and each staged G row is read once from smem.
  const int npairs = (nrows + 1) >> 1;
  for (int c0 = 0; c0 < p.C; c0 += kGT) {
    const int ct = min(kGT, p.C - c0);
    __syncthreads();
    for (int i = tid; i < ct * D4; i += kThreads) st4(s_G + 4 * (size_t)i, ld4(p.ws_g + (size_t)c0 * D + 4 * (size_t)i));
    __syncthreads();
    for (int pr = warp; pr < npairs; pr += kWarps) {
      const int lr0 = 2 * pr, lr1 = min(2 * pr + 1, nrows - 1);
      float acc0[kGT], acc1[kGT];
#pragma unroll
      for (int c = 0; c < kGT; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
      for (int jb = 0; jb < D; jb += kChunk) {
        float4 a0[8], a1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = jb + (q * 32 + lane) * 4;
          if (j < D) {
            a0[q] = ld4(s_P + (size_t)lr0 * D + j);
            a1[q] = ld4(s_P + (size_t)lr1 * D + j);
          } else {
            a0[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            a1[q] = a0[q];
          }
        }
#pragma unroll
        for (int c = 0; c < kGT; ++c) {
          if (c < ct) {
            float s0 = acc0[c], s1 = acc1[c];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = jb + (q * 32 + lane) * 4;
              if (j < D) {
                const float4 g = ld4(s_G + (size_t)c * D + j);
                s0 = fmaf(g.x, a0[q].x, s0); s0 = fmaf(g.y, a0[q].y, s0);
                s0 = fmaf(g.z, a0[q].z, s0); s0 = fmaf(g.w, a0[q].w, s0);
                s1 = fmaf(g.x, a1[q].x, s1); s1 = fmaf(g.y, a1[q].y, s1);
                s1 = fmaf(g.z, a1[q].z, s1); s1 = fmaf(g.w, a1[q].w, s1);
              }
            }
            acc0[c] = s0; acc1[c] = s1;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kGT; ++c) {
        if (c < ct) {
          const float s0 = mla::warp_sum(acc0[c]);
          const float s1 = mla::warp_sum(acc1[c]);
          if (lane == 0) {
            p.grad_w[(size_t)(c0 + c) * D + row0 + lr0] = s0;
            if (lr1 != lr0) p.grad_w[(size_t)(c0 + c) * D + row0 + lr1] = s1;
          }
        }
      }
    }
  }
}

struct GsPlan {
  int grid, rows_per_cta, nb, rows_per_nb;
  size_t smem;
  size_t off_r, off_k, off_part, off_norm, off_g, total;
};

int make_plan(int B, int D, int C, GsPlan* pl) {
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (B < 1 || D < 4 || (D & 3) || C < 0 || C > 4096) return MLA_E_SHAPE;
  const int sms = di.sm_count;
  int rpc = (D + sms - 1) / sms;
  if (rpc < 4) rpc = min(4, D);          // tiny D: fewer, fuller CTAs
  int grid = (D + rpc - 1) / rpc;
  size_t smem = ((size_t)2 * D + (size_t)rpc * D + (size_t)kGT * D + 32) * sizeof(float);
  if (smem > (size_t)di.smem_optin) return MLA_E_SHAPE;
  // raw-feature reduction: >= 8 rows per unit, about one unit per warp of the grid
  const int cchunks = (D + 127) / 128;
  int nb = (grid * kWarps) / cchunks;
  nb = max(1, min(nb, (B + 7) / 8));
  int rows_per_nb = (B + nb - 1) / nb;
  nb = (B + rows_per_nb - 1) / rows_per_nb;
  pl->grid = grid; pl->rows_per_cta = rpc; pl->nb = nb; pl->rows_per_nb = rows_per_nb; pl->smem = smem;
  size_t off = 0;
  pl->off_r = off;    off += mla::align_up((size_t)D * 4, 256);
  pl->off_k = off;    off += mla::align_up((size_t)D * 4, 256);
  pl->off_part = off; off += mla::align_up((size_t)nb * D * 4, 256);
  pl->off_norm = off; off += mla::align_up((size_t)grid * 8, 256);
  pl->off_g = off;    off += mla::align_up((size_t)max(C, 1) * D * 4, 256);
  pl->total = off;
  return 0;
}

}  // namespace

extern "C" size_t mla_gs_project_workspace_bytes(int B, int D, int C) {
  GsPlan pl;
  if (make_plan(B, D, C, &pl) != 0) return 0;
  return pl.total;
}

extern "C" int mla_gs_project(float* P, const float* feat, const float* feat_sum, float inv_batch,
                              float alpha, float* grad_w, int B, int D, int C, int mode,
                              void* ws, size_t ws_bytes, void* stream) {
  if (P == nullptr || ((feat == nullptr) == (feat_sum == nullptr))) return MLA_E_BADARG;
  if (mode != 0 && mode != 1) return MLA_E_BADARG;
  if (!mla::aligned16(P) || !mla::aligned16(feat) || !mla::aligned16(feat_sum) || !mla::aligned16(grad_w) ||
      !mla::aligned16(ws))
    return MLA_E_BADARG;
  if (grad_w == nullptr) C = 0;
  GsPlan pl;
  int rc = make_plan(B, D, C, &pl);
  if (rc != 0) return rc;
  if (ws == nullptr || ws_bytes < pl.total) return MLA_E_WORKSPACE;
  const mla::DeviceInfo& di = mla::device_info();
  if (!di.coop) return MLA_E_NODEVICE;

  static std::atomic<size_t> s_smem_set{0};
  if (pl.smem > s_smem_set.load(std::memory_order_relaxed)) {
    MLA_CUDA_TRY(cudaFuncSetAttribute(gs_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)di.smem_optin));
    s_smem_set.store((size_t)di.smem_optin, std::memory_order_relaxed);
  }
  char* w = static_cast<char*>(ws);
  GsParams prm;
  prm.P = P; prm.feat = feat; prm.feat_sum = feat_sum; prm.inv_batch = inv_batch; prm.alpha = alpha;
  prm.grad_w = grad_w; prm.B = B; prm.D = D; prm.C = C; prm.mode = mode;
  prm.rows_per_cta = pl.rows_per_cta; prm.nb = pl.nb; prm.rows_per_nb = pl.rows_per_nb;
  prm.ws_r = reinterpret_cast<float*>(w + pl.off_r);
  prm.ws_k = reinterpret_cast<float*>(w + pl.off_k);
  prm.ws_part = reinterpret_cast<float*>(w + pl.off_part);
  prm.ws_norm = reinterpret_cast<double*>(w + pl.off_norm);
  prm.ws_g = reinterpret_cast<float*>(w + pl.off_g);
  void* args[] = {&prm};
  MLA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)gs_project_kernel, dim3(pl.grid), dim3(kThreads), args,
                                           pl.smem, static_cast<cudaStream_t>(stream)));
  mla::count_launch();
  return 0;
}
