// Fused GSPlugin projection — replaces the body of GSPlugin.before_update
// (reference utils/utils.py:34-41). One cooperative launch:
//
//   phase 0  one thread issues bulk async copies (cp.async.bulk, completing on an mbarrier) of the CTA's row slice of P
//            into shared memory: P leaves HBM ONCE, with zero register cost, while every warp of the CTA streams
//            its share of feat (raw-feature path: CTA-blocked rows, up to 16 independent 512-byte loads in flight per
//            warp, fixed-order partial sums) and grad_w is copied to the workspace (the in-place projection can't race)
//   phase 1  r = inv_batch * sum of the per-CTA partials (fixed order), one column slice per thread of the grid
//   phase A  k = P r^T          from the shared-memory rows
//   phase B  P' = P - (k k^T) ./ (alpha + k r)   in shared memory, + sum(P'^2) partials
//   phase C  P = P' / ||P'||_F  written to HBM once; grad_w = grad_w @ P^T from the smem-resident rows: every warp owns a
//            (4-row group, 512-column chunk) block held in registers, staged gradient rows are read once per 16 FMAs
//
// HBM traffic is the algorithmic minimum 4*(B*D + 2*D*D + 2*C*D) bytes: feat is read once, P is read once and
// written once. All reductions have a fixed order, so every rank of a data-parallel job
// that feeds identical (P, feat_sum, grad_w) computes bit-identical results.
#include "common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kGT = 8;        // grad rows staged per tile in phase C
constexpr int kRB = 4;        // P rows per register block in phase C
constexpr int kCB = 512;      // columns per register block in phase C (4 float4 per lane and row)
constexpr int kMaxCB = 4;     // D <= kMaxCB * kCB

struct GsParams {
  float* P;
  const float* feat;
  const float* feat_sum;
  float inv_batch;
  float alpha;
  float* grad_w;
  int B, D, C, mode;
  int rows_per_cta;
  int nb;           // CTAs that hold a batch slice in the raw-feature reduction
  int rows_per_nb;  // batch rows per such CTA
  float* ws_r;      // [D]
  float* ws_k;      // [D]
  float* ws_part;   // [nb][D]
  double* ws_norm;  // [grid]
  float* ws_g;      // [C][D]
  unsigned long long* ws_ts;   // [8] %globaltimer at the phase boundaries of CTA 0 (profiling aid)
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) gs_project_kernel(GsParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t s_bar;
  const int D = p.D, D4 = p.D >> 2;
  float* s_r = smem;                      // [D]
  float* s_k = s_r + D;                   // [D]
  float* s_P = s_k + D;                   // [rows_per_cta][D]
  float* s_G = s_P + (size_t)p.rows_per_cta * D;  // [kGT][D]   (phase 0: scratch of the feature reduction)
  float* s_red = s_G + (size_t)kGT * D;   // [kWarps * kGT * kRB]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gtid = blockIdx.x * kThreads + tid;
  const int gthreads = gridDim.x * kThreads;
  const int row0 = blockIdx.x * p.rows_per_cta;
  const int nrows = max(0, min(p.rows_per_cta, D - row0));
  const bool stamp = (blockIdx.x == 0 && tid == 0);
  if (stamp) p.ws_ts[0] = tc::globaltimer_ns();
  if (tid == 0) p.ws_ts[176 + blockIdx.x] = tc::globaltimer_ns();     // every CTA's start

  // ---------------- phase 0: P rows -> smem (async), feat partial sums, grad copy
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&s_bar), 1);
    tc::fence_mbar_init();
    if (nrows > 0) {
      tc::mbar_arrive_expect_tx(tc::smem_u32(&s_bar), (uint32_t)nrows * (uint32_t)D * 4u);
      for (int lr = 0; lr < nrows; ++lr)
        bulk_g2s(tc::smem_u32(s_P + (size_t)lr * D), p.P + (size_t)(row0 + lr) * D, (uint32_t)D * 4u, tc::smem_u32(&s_bar));
    }
  }
  if (p.grad_w != nullptr) {
    const int n4 = p.C * D4;
    for (int i = gtid; i < n4; i += gthreads) st4(p.ws_g + 4 * (size_t)i, ld4(p.grad_w + 4 * (size_t)i));
  }
  if (p.feat != nullptr) {
    // CTA c < nb sums batch rows [c * rows_per_nb, ...). Inside the CTA a unit = (128-column chunk, row split): with few
    // column chunks (small D) the rows are split over the warps as well; the row-split partials are combined through
    // shared memory in split order (fixed -> deterministic).
    if ((int)blockIdx.x < p.nb) {
      const int cchunks = (D + 127) / 128;
      const int rsplit = min(kGT, max(1, kWarps / cchunks));     // the scratch holds kGT rows of D
      const int b0 = blockIdx.x * p.rows_per_nb, b1 = min(p.B, b0 + p.rows_per_nb);
      float* scratch = s_G;                                     // [rsplit][D]  (rsplit * D <= kGT * D)
      for (int u = warp; u < cchunks * rsplit; u += kWarps) {
        const int cc = u % cchunks, rs = u / cchunks;
        const int col = cc * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < D) {
          const float* src = p.feat + col;
          int b = b0 + rs;
          for (; b + 15 * rsplit < b1; b += 16 * rsplit) {      // 16 independent row segments in flight
            float4 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = ldcs4(src + (size_t)(b + q * rsplit) * D);
#pragma unroll
            for (int q = 0; q < 16; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
          }
          {                                                     // tail: the remaining (< 16) rows, all loads first
            float4 v[16];
            int n = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (b + q * rsplit < b1) { v[q] = ldcs4(src + (size_t)(b + q * rsplit) * D); n = q + 1; }
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (q < n) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
          }
          st4(scratch + (size_t)rs * D + col, acc);
        }
      }
      __syncthreads();
      for (int j4 = tid; j4 < D4; j4 += kThreads) {
        float4 s = ld4(scratch + 4 * j4);
        for (int rs = 1; rs < rsplit; ++rs) {
          const float4 v = ld4(scratch + (size_t)rs * D + 4 * j4);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        st4(p.ws_part + (size_t)blockIdx.x * D + 4 * j4, s);
      }
    }
    if (stamp) p.ws_ts[1] = tc::globaltimer_ns();
    grid.sync();
    if (stamp) p.ws_ts[2] = tc::globaltimer_ns();
    // r[j] = inv_batch * sum_c part[c][j]. Eight lanes share a column: lane q adds the partials c = q, q + 8, ... (all of its
    // loads in flight at once), then the eight sub-sums are added in lane order — a fixed tree, identical on every rank.
    {
      const int q = lane & 7;
      for (int j = (gtid >> 3); j < D; j += (gthreads >> 3)) {
        float v[24];
        int n = 0;
#pragma unroll
        for (int u = 0; u < 24; ++u) {
          const int c = q + 8 * u;
          if (c < p.nb) { v[u] = __ldcg(p.ws_part + (size_t)c * D + j); n = u + 1; }
        }
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 24; ++u)
          if (u < n) s += v[u];
        for (int c = q + 8 * 24; c < p.nb; c += 8) s += __ldcg(p.ws_part + (size_t)c * D + j);      // nb > 192: not on 148 SMs
        // lanes 8k .. 8k+7 hold the sub-sums of one column: add them in lane order
        float t = __shfl_sync(0xffffffffu, s, (lane & ~7));
#pragma unroll
        for (int u = 1; u < 8; ++u) t += __shfl_sync(0xffffffffu, s, (lane & ~7) + u);
        if (q == 0) p.ws_r[j] = t * p.inv_batch;
      }
    }
    grid.sync();
    for (int j = tid; j < D; j += kThreads) s_r[j] = __ldcg(p.ws_r + j);
  } else {
    for (int j = tid; j < D; j += kThreads) s_r[j] = p.feat_sum[j] * p.inv_batch;
  }
  __syncthreads();
  if (stamp) p.ws_ts[3] = tc::globaltimer_ns();

  // ---------------- phase A: k_i = sum_j P_ij r_j from the smem-resident rows
  if (nrows > 0) tc::mbar_wait(tc::smem_u32(&s_bar), 0);
  for (int lr = warp; lr < nrows; lr += kWarps) {
    const float* src = s_P + (size_t)lr * D;
    float acc = 0.f;
    for (int j4 = lane; j4 < D4; j4 += 32) {
      const float4 v = ld4(src + 4 * j4);
      const float4 r = ld4(s_r + 4 * j4);
      acc = fmaf(v.x, r.x, acc); acc = fmaf(v.y, r.y, acc);
      acc = fmaf(v.z, r.z, acc); acc = fmaf(v.w, r.w, acc);
    }
    acc = mla::warp_sum(acc);
    if (lane == 0) p.ws_k[row0 + lr] = acc;
  }
  if (stamp) p.ws_ts[4] = tc::globaltimer_ns();
  grid.sync();

  // ---------------- phase B: elementwise update in smem + sum of squares
  for (int j = tid; j < D; j += kThreads) s_k[j] = __ldcg(p.ws_k + j);
  __syncthreads();
  float scal_den = 0.f;
  if (p.mode == 1) {  // canonical OWM: scalar denominator alpha + r.k (same order in every CTA)
    float part = 0.f;
    for (int j = tid; j < D; j += kThreads) part = fmaf(s_r[j], s_k[j], part);
    scal_den = __fadd_rn(p.alpha, mla::block_sum(part, s_red));
  }
  float sq = 0.f;
  for (int lr = 0; lr < nrows; ++lr) {
    const float ki = s_k[row0 + lr];
    float* prow = s_P + (size_t)lr * D;
    for (int j4 = tid; j4 < D4; j4 += kThreads) {
      float4 pv = ld4(prow + 4 * j4);
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
      st4(prow + 4 * j4, pv);
    }
  }
  {
    const float bs = mla::block_sum(sq, s_red);
    if (tid == 0) p.ws_norm[blockIdx.x] = (double)bs;
  }
  if (stamp) p.ws_ts[5] = tc::globaltimer_ns();
  if (tid == 0) p.ws_ts[16 + blockIdx.x] = tc::globaltimer_ns();     // every CTA's arrival at the last barrier
  grid.sync();
  if (stamp) p.ws_ts[8] = tc::globaltimer_ns();

  // ---------------- phase C: normalise, write P, project the gradient
  // ||P'||_F^2 = sum of the per-CTA partials in CTA order (identical in every CTA): the partials are fetched once per CTA
  // (one L2 load per thread, NOT one per thread and partial: 75k threads polling the same 148 lines cost 30 us) and added
  // by one thread from shared memory
  {
    double* s_nrm = reinterpret_cast<double*>(s_G);               // s_G is free until the projection
    const int G = (int)gridDim.x;
    for (int c = tid; c < G; c += kThreads) s_nrm[c] = __ldcg(p.ws_norm + c);
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int c = 0; c < G; ++c) tot += s_nrm[c];
      s_red[0] = (float)sqrt(tot);
    }
    __syncthreads();
  }
  const float nrm = s_red[0];
  if (stamp) p.ws_ts[9] = tc::globaltimer_ns();
  for (int lr = 0; lr < nrows; ++lr) {
    float* prow = s_P + (size_t)lr * D;
    float4* grow = reinterpret_cast<float4*>(p.P + (size_t)(row0 + lr) * D);
    for (int j4 = tid; j4 < D4; j4 += kThreads) {
      float4 pv = ld4(prow + 4 * j4);
      pv.x = __fdiv_rn(pv.x, nrm); pv.y = __fdiv_rn(pv.y, nrm);
      pv.z = __fdiv_rn(pv.z, nrm); pv.w = __fdiv_rn(pv.w, nrm);
      st4(prow + 4 * j4, pv);
      __stcs(grow + j4, pv);
    }
  }
  if (stamp) p.ws_ts[6] = tc::globaltimer_ns();
  if (p.grad_w == nullptr) {
    if (tid == 0) p.ws_ts[336 + blockIdx.x] = tc::globaltimer_ns();
    return;
  }
  __syncthreads();

  // grad_w[c][i] = sum_j G[c][j] * P[i][j]. Work unit = (group of kRB rows, kCB-column chunk): the unit's P block sits in
  // registers (kRB x 4 float4 per lane) and every staged G value is loaded once per kRB * 4 FMAs. A pass covers whole row
  // groups (all their column chunks); when a pass has fewer units than warps, several warps share a unit and split the
  // staged gradient rows. The column-chunk partials of one (c, row) are combined through shared memory in chunk order.
  const int ngroups = (nrows + kRB - 1) / kRB;
  const int nchunks = (D + kCB - 1) / kCB;
  const int gpp = max(1, kWarps / nchunks);                     // row groups per pass
  for (int c0 = 0; c0 < p.C; c0 += kGT) {
    const int ct = min(kGT, p.C - c0);
    __syncthreads();
    for (int i = tid; i < ct * D4; i += kThreads)
      st4(s_G + 4 * (size_t)i, __ldcg(reinterpret_cast<const float4*>(p.ws_g + (size_t)c0 * D) + i));
    __syncthreads();
    for (int g0 = 0; g0 < ngroups; g0 += gpp) {
      const int gn = min(gpp, ngroups - g0);                    // groups in this pass
      const int units = gn * nchunks;                           // <= kWarps
      const int csplit = max(1, kWarps / units);
      const int ul = warp % units, csub = warp / units;
      if (csub < csplit) {
        const int g = g0 + ul / nchunks, ch = ul % nchunks;
        const int jb = ch * kCB;
        float4 a[kRB][4];
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
          const int lr = min(g * kRB + rr, nrows - 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = jb + (q * 32 + lane) * 4;
            a[rr][q] = (j < D) ? ld4(s_P + (size_t)lr * D + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        for (int c = csub; c < ct; c += csplit) {
          float sacc[kRB];
#pragma unroll
          for (int rr = 0; rr < kRB; ++rr) sacc[rr] = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = jb + (q * 32 + lane) * 4;
            if (j < D) {
              const float4 gv = ld4(s_G + (size_t)c * D + j);
#pragma unroll
              for (int rr = 0; rr < kRB; ++rr) {
                sacc[rr] = fmaf(gv.x, a[rr][q].x, sacc[rr]); sacc[rr] = fmaf(gv.y, a[rr][q].y, sacc[rr]);
                sacc[rr] = fmaf(gv.z, a[rr][q].z, sacc[rr]); sacc[rr] = fmaf(gv.w, a[rr][q].w, sacc[rr]);
              }
            }
          }
#pragma unroll
          for (int rr = 0; rr < kRB; ++rr) {
            const float t = mla::warp_sum(sacc[rr]);
            if (lane == 0) s_red[(ul * kGT + c) * kRB + rr] = t;
          }
        }
      }
      __syncthreads();
      for (int t = tid; t < gn * ct * kRB; t += kThreads) {      // (group, c, row): chunks added in chunk order
        const int rr = t % kRB, c = (t / kRB) % ct, gl = t / (kRB * ct);
        const int lr = (g0 + gl) * kRB + rr;
        if (lr < nrows) {
          float acc = 0.f;
          for (int ch = 0; ch < nchunks; ++ch) acc += s_red[((gl * nchunks + ch) * kGT + c) * kRB + rr];
          p.grad_w[(size_t)(c0 + c) * D + row0 + lr] = acc;
        }
      }
      __syncthreads();
    }
  }
  if (stamp) p.ws_ts[7] = tc::globaltimer_ns();
  if (tid == 0) p.ws_ts[336 + blockIdx.x] = tc::globaltimer_ns();     // every CTA's end
}

struct GsPlan {
  int grid, rows_per_cta, nb, rows_per_nb;
  size_t smem;
  size_t off_r, off_k, off_part, off_norm, off_g, off_ts, total;
};

int make_plan(int B, int D, int C, GsPlan* pl) {
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  if (B < 1 || D < 4 || (D & 3) || C < 0 || C > 4096) return MLA_E_SHAPE;
  const int sms = di.sm_count;
  int rpc = (D + sms - 1) / sms;
  if (rpc < 4) rpc = min(4, D);          // tiny D: fewer, fuller CTAs
  int grid = (D + rpc - 1) / rpc;
  if (D > kMaxCB * kCB) return MLA_E_SHAPE;
  size_t smem = ((size_t)2 * D + (size_t)rpc * D + (size_t)kGT * D + (size_t)kWarps * kGT * kRB + 32) * sizeof(float);
  if (smem + 256 > (size_t)di.smem_optin) return MLA_E_SHAPE;
  // raw-feature reduction: the batch rows are dealt to the CTAs in contiguous slices (>= 1 row each); inside a CTA the
  // warps split columns (and rows, when D has fewer than kWarps 128-column chunks)
  int rows_per_nb = (B + grid - 1) / grid;
  int nb = (B + rows_per_nb - 1) / rows_per_nb;
  pl->grid = grid; pl->rows_per_cta = rpc; pl->nb = nb; pl->rows_per_nb = rows_per_nb; pl->smem = smem;
  size_t off = 0;
  pl->off_r = off;    off += mla::align_up((size_t)D * 4, 256);
  pl->off_k = off;    off += mla::align_up((size_t)D * 4, 256);
  pl->off_part = off; off += mla::align_up((size_t)nb * D * 4, 256);
  pl->off_norm = off; off += mla::align_up((size_t)grid * 8, 256);
  pl->off_g = off;    off += mla::align_up((size_t)max(C, 1) * D * 4, 256);
  pl->off_ts = off;   off += 4096;
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
    // the opt-in limit covers static + dynamic shared memory: leave room for the kernel's static barrier word
    const size_t want = std::min((size_t)di.smem_optin - 256, std::max(pl.smem, (size_t)128 * 1024));
    MLA_CUDA_TRY(cudaFuncSetAttribute(gs_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    s_smem_set.store(want, std::memory_order_relaxed);
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
  prm.ws_ts = reinterpret_cast<unsigned long long*>(w + pl.off_ts);
  void* args[] = {&prm};
  MLA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)gs_project_kernel, dim3(pl.grid), dim3(kThreads), args,
                                           pl.smem, static_cast<cudaStream_t>(stream)));
  mla::count_launch();
  return 0;
}
