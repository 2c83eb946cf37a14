// Fused GSPlugin projection — replaces the body of GSPlugin.before_update
// (reference utils/utils.py:34-41). One cooperative launch:
//
//   phase 0  one thread issues bulk async copies (cp.async.bulk, completing on an mbarrier) of the CTA's row slice of P
//            into shared memory: P leaves HBM ONCE, with zero register cost, while every warp of the CTA streams
//            its share of feat (raw-feature path: CTA = (batch slice, 256-column chunk), up to 16 independent 512-byte
//            loads in flight per warp, fixed-order partial sums) and grad_w is copied to the workspace (the in-place
//            projection can't race)
//   phase 1  r = inv_batch * sum of the <= 20 partial rows (slice order), formed by every CTA for itself: one grid barrier
//   phase A  k = P r^T          from the shared-memory rows
//   phase B  P' = P - (k k^T) ./ (alpha + k r)   in shared memory, + sum(P'^2) partials
//   phase C  P = P' / ||P'||_F  written to HBM once; grad_w = grad_w @ P^T from the smem-resident rows: a warp owns a
//            (4-row group, column chunk) block of P in registers and walks the gradient rows four at a time; the gradient is
//            streamed through two cp.async-fed staging buffers (tile t+1 lands while tile t is multiplied, one barrier per
//            tile), the 16 lane-partial sums of a (4 classes x 4 rows) block are reduced with one butterfly transpose
//            (16 shuffles instead of 80), column-chunk partials are combined in chunk order
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
constexpr int kRB = 4;        // P rows per register block in phase C
constexpr int kCT = 4;        // gradient rows (classes) per register block in phase C
constexpr int kMaxD = 2048;   // widest P (14 rows per CTA on 148 SMs = 112 KB of shared memory)
constexpr int kStageFloats = 8192;   // one gradient staging buffer (32 KB); there are two
constexpr int kMaxStageRows = 16;
constexpr int kMaxUnits = 64;        // (row group, column chunk, class block) units of one staged tile

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
  int stage_rows;   // gradient rows per staging buffer (multiple of kCT)
  int nq;           // float4 per lane of a projection column chunk (chunk = nq * 128 columns): 1, 2 or 4
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

// One (4-row group, nq * 128-column chunk, 4-class block) unit of the projection. `a` = the unit's block of P (registers,
// loaded by the caller); G = staged gradient rows [ct][D]. The 16 lane-partial sums acc[class][row] are reduced over the
// warp with a butterfly transpose: after the four exchange steps lane l holds the sum (over one half-warp) of value l & 15,
// the last step adds the two half-warps. Fixed order -> deterministic. Lanes 0..15 store value `lane` to out[lane].
template <int NQ>
__device__ __forceinline__ void project_unit(const float4 (&a)[kRB][4], const float* __restrict__ G, int D, int jb, int c_lo,
                                             int ct, int lane, float* __restrict__ out) {
  float acc[kCT * kRB];
#pragma unroll
  for (int i = 0; i < kCT * kRB; ++i) acc[i] = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int j = jb + (q * 32 + lane) * 4;
    if (j < D) {
      float4 gv[kCT];
#pragma unroll
      for (int cc = 0; cc < kCT; ++cc) {
        const int c = min(c_lo + cc, ct - 1);                      // rows past the tile repeat the last one (never stored)
        gv[cc] = ld4(G + (size_t)c * D + j);
      }
#pragma unroll
      for (int cc = 0; cc < kCT; ++cc) {
#pragma unroll
        for (int rr = 0; rr < kRB; ++rr) {
          float t = acc[cc * kRB + rr];
          t = fmaf(gv[cc].x, a[rr][q].x, t); t = fmaf(gv[cc].y, a[rr][q].y, t);
          t = fmaf(gv[cc].z, a[rr][q].z, t); t = fmaf(gv[cc].w, a[rr][q].w, t);
          acc[cc * kRB + rr] = t;
        }
      }
    }
  }
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? acc[i] : acc[i + o];
      const float keep = up ? acc[i + o] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  const float tot = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 16);
  if (lane < 16) out[lane] = tot;
}

__global__ void __launch_bounds__(kThreads, 1) gs_project_kernel(GsParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float s_nrmf;
  const int D = p.D, D4 = p.D >> 2;
  float* s_r = smem;                      // [D]
  float* s_k = s_r + D;                   // [D]
  float* s_P = s_k + D;                   // [rows_per_cta][D]
  float* s_G = s_P + (size_t)p.rows_per_cta * D;  // [2][kStageFloats] gradient staging (phase 0: feature-reduction scratch)
  float* s_red = s_G + 2 * kStageFloats;  // [2][kMaxUnits][16] projection partials; block_sum scratch; norm partials

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
    // Feature reduction. CTA (s, ch) = (batch slice, 256-column chunk) sums rows [s * rows_per_nb, ...) of its columns: warp
    // = (128-column half, one of 8 row splits), up to 16 independent 512-byte row segments in flight per warp; the row-split
    // partials are combined through shared memory in split order. Slicing by columns as well keeps the number of partial
    // rows small (nb <= 20 at D = 2048), so that after ONE grid barrier every CTA can form the whole of r by itself.
    const int cch = (D + 255) / 256;
    if ((int)blockIdx.x < p.nb * cch) {
      const int sl = blockIdx.x / cch, ch = blockIdx.x - sl * cch;
      const int b0 = sl * p.rows_per_nb, b1 = min(p.B, b0 + p.rows_per_nb);
      const int h = warp & 1, rs = warp >> 1;
      constexpr int kRS = kWarps / 2;
      float* scratch = s_G;                                     // [kRS][256]
      const int col = ch * 256 + h * 128 + lane * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < D) {
        const float* src = p.feat + col;
        int b = b0 + rs;
        for (; b + 15 * kRS < b1; b += 16 * kRS) {              // 16 independent row segments in flight
          float4 v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = ldcs4(src + (size_t)(b + q * kRS) * D);
#pragma unroll
          for (int q = 0; q < 16; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        }
        {                                                       // tail: the remaining (< 16) rows, all loads first
          float4 v[16];
          int n = 0;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (b + q * kRS < b1) { v[q] = ldcs4(src + (size_t)(b + q * kRS) * D); n = q + 1; }
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < n) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        }
      }
      st4(scratch + rs * 256 + h * 128 + lane * 4, acc);
      __syncthreads();
      if (tid < 64 && ch * 256 + 4 * tid < D) {
        float4 t = ld4(scratch + 4 * tid);
#pragma unroll
        for (int q = 1; q < kRS; ++q) {
          const float4 v = ld4(scratch + q * 256 + 4 * tid);
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(p.ws_part + (size_t)sl * D + ch * 256 + 4 * tid, t);
      }
    }
    if (stamp) p.ws_ts[1] = tc::globaltimer_ns();
    grid.sync();
    if (stamp) p.ws_ts[2] = tc::globaltimer_ns();
    // r = inv_batch * (sum of the nb partial rows, in slice order): formed by every CTA for itself, eight loads in flight
    for (int j4 = tid; j4 < D4; j4 += kThreads) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* src = reinterpret_cast<const float4*>(p.ws_part) + j4;
      int c = 0;
      for (; c + 8 <= p.nb; c += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldcg(src + (size_t)(c + q) * D4);
#pragma unroll
        for (int q = 0; q < 8; ++q) { t.x += v[q].x; t.y += v[q].y; t.z += v[q].z; t.w += v[q].w; }
      }
      for (; c < p.nb; ++c) {
        const float4 v = __ldcg(src + (size_t)c * D4);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      t.x *= p.inv_batch; t.y *= p.inv_batch; t.z *= p.inv_batch; t.w *= p.inv_batch;
      st4(s_r + 4 * j4, t);
    }
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

  // the workspace copy of the gradient is complete (grid barrier above): start streaming its first tile now, it lands while
  // the update below runs. Staging buffer b holds gradient rows [t * stage_rows, ...) of tile t, t & 1 == b.
  const int ntiles = (p.grad_w != nullptr && nrows > 0) ? (p.C + p.stage_rows - 1) / p.stage_rows : 0;
  auto issue_tile = [&](int t) {
    const int c0 = t * p.stage_rows, ct = min(p.stage_rows, p.C - c0);
    const float* src = p.ws_g + (size_t)c0 * D;
    const uint32_t dst = tc::smem_u32(s_G + (size_t)(t & 1) * kStageFloats);
    for (int i = tid; i < ct * D4; i += kThreads) tc::cp_async16(dst + 16u * (uint32_t)i, src + 4 * (size_t)i, 16u);
    tc::cp_async_commit();
  };
  if (ntiles > 0) issue_tile(0);
  if (ntiles > 1) issue_tile(1);                                // both staging buffers are free

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
  // a thread owns float4 column j4 of every row: k_j / r_j are read once, four rows are updated together (16 independent
  // divisions in flight)
  for (int j4 = tid; j4 < D4; j4 += kThreads) {
    const float4 kv = ld4(s_k + 4 * j4);
    const float4 rv = ld4(s_r + 4 * j4);
    for (int lr0 = 0; lr0 < nrows; lr0 += 4) {
      float4 pv[4];
      float ki[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int lr = min(lr0 + u, nrows - 1);
        pv[u] = ld4(s_P + (size_t)lr * D + 4 * j4);
        ki[u] = s_k[row0 + lr];
      }
      // Same operation order and roundings as utils.py:36 (no FMA contraction):
      //   P - (k_i*k_j) / (alpha + k_i*r_j)
#define MLA_GS_UPD(u, c)                                                                       \
      {                                                                                        \
        const float den = (p.mode == 0) ? __fadd_rn(p.alpha, __fmul_rn(ki[u], rv.c)) : scal_den; \
        pv[u].c = __fsub_rn(pv[u].c, __fdiv_rn(__fmul_rn(ki[u], kv.c), den));                  \
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { MLA_GS_UPD(u, x) MLA_GS_UPD(u, y) MLA_GS_UPD(u, z) MLA_GS_UPD(u, w) }
#undef MLA_GS_UPD
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (lr0 + u < nrows) {
          st4(s_P + (size_t)(lr0 + u) * D + 4 * j4, pv[u]);
          sq = fmaf(pv[u].x, pv[u].x, sq); sq = fmaf(pv[u].y, pv[u].y, sq);
          sq = fmaf(pv[u].z, pv[u].z, sq); sq = fmaf(pv[u].w, pv[u].w, sq);
        }
      }
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
  // ||P'||_F^2 = sum of the per-CTA partials, identical in every CTA: the partials are fetched once per CTA (one L2 load per
  // thread), lane l of warp 0 adds partials l, l + 32, ... in order, then a fixed shuffle tree (fp64)
  {
    double* s_nrm = reinterpret_cast<double*>(s_red);
    const int G = (int)gridDim.x;
    for (int c = tid; c < G; c += kThreads) s_nrm[c] = __ldcg(p.ws_norm + c);
    __syncthreads();
    if (warp == 0) {
      double tot = 0.0;
      for (int c = lane; c < G; c += 32) tot += s_nrm[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0) s_nrmf = (float)sqrt(tot);
    }
    __syncthreads();
  }
  const float nrm = s_nrmf;
  if (stamp) p.ws_ts[9] = tc::globaltimer_ns();
  for (int j4 = tid; j4 < D4; j4 += kThreads) {
    for (int lr0 = 0; lr0 < nrows; lr0 += 4) {
      float4 pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pv[u] = ld4(s_P + (size_t)min(lr0 + u, nrows - 1) * D + 4 * j4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pv[u].x = __fdiv_rn(pv[u].x, nrm); pv[u].y = __fdiv_rn(pv[u].y, nrm);
        pv[u].z = __fdiv_rn(pv[u].z, nrm); pv[u].w = __fdiv_rn(pv[u].w, nrm);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (lr0 + u < nrows) {
          st4(s_P + (size_t)(lr0 + u) * D + 4 * j4, pv[u]);
          __stcs(reinterpret_cast<float4*>(p.P + (size_t)(row0 + lr0 + u) * D) + j4, pv[u]);
        }
      }
    }
  }
  // (Measured alternatives that did NOT help this phase and the projection after it at D = 2048: writing P back with bulk
  // shared -> global copies issued by one thread instead of these streaming stores (+2 us); a rolled, shared-memory-fed
  // projection loop with a third of the code size (equal at C = 6, +3 us at C = 101); component-outermost FMA order in the
  // projection (+2 us). The normalisation loop itself — 28 k IEEE divisions per SM at 16 warps — is what bounds the phase.)
  if (stamp) p.ws_ts[15] = tc::globaltimer_ns();
  if (stamp) p.ws_ts[6] = tc::globaltimer_ns();
  if (ntiles == 0) {
    if (tid == 0) p.ws_ts[336 + blockIdx.x] = tc::globaltimer_ns();
    return;
  }

  // grad_w[c][i] = sum_j G[c][j] * P[i][j] for the CTA's rows i. Unit = (group g of kRB rows, column chunk ch, class block
  // cb of kCT staged rows); index ((cb * ngroups) + g) * nchunks + ch. A warp keeps one (g, ch) block of P in registers for as
  // long as it can (always, when there are at most kWarps such blocks) and walks the class blocks of every tile.
  const int ngroups = (nrows + kRB - 1) / kRB;
  const int chunk = p.nq * 128;
  const int nchunks = (D + chunk - 1) / chunk;
  const int GC = ngroups * nchunks;
  const int wpg = max(1, kWarps / GC);                          // warps sharing a (g, ch) block (they split the class blocks)
  float4 a[kRB][4];
  int have_gc = -1;
  auto load_block = [&](int gc) {
    if (gc == have_gc) return;
    have_gc = gc;
    const int g = gc / nchunks, jb = (gc - g * nchunks) * chunk;
#pragma unroll
    for (int rr = 0; rr < kRB; ++rr) {
      const int lr = min(g * kRB + rr, nrows - 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb + (q * 32 + lane) * 4;
        a[rr][q] = (q < p.nq && j < D) ? ld4(s_P + (size_t)lr * D + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  for (int t = 0; t <= ntiles; ++t) {
    if (t == 0 && ntiles > 1) tc::cp_async_wait<1>(); else tc::cp_async_wait<0>();   // tile t has landed (this thread's part) ...
    __syncthreads();                // ... and everybody's; every warp has finished tile t - 1 (its partials are complete)
    if (stamp && t < 3) p.ws_ts[10 + 2 * t] = tc::globaltimer_ns();
    if (t >= 1 && t + 1 < ntiles) issue_tile(t + 1);             // into the buffer tile t - 1 has just released
    if (t >= 1) {                   // combine tile t - 1: column-chunk partials in chunk order
      const int tp = t - 1, c0 = tp * p.stage_rows, ct = min(p.stage_rows, p.C - c0);
      const float* red = s_red + (size_t)(tp & 1) * kMaxUnits * 16;
      for (int o = tid; o < ct * ngroups * kRB; o += kThreads) {
        const int rr = o % kRB, g = (o / kRB) % ngroups, c = o / (kRB * ngroups);
        const int lr = g * kRB + rr;
        if (lr < nrows) {
          const int cb = c / kCT, cc = c - cb * kCT;
          const float* src = red + ((size_t)(cb * ngroups + g) * nchunks) * 16 + cc * kRB + rr;
          float acc = 0.f;
          for (int ch = 0; ch < nchunks; ++ch) acc += src[ch * 16];
          p.grad_w[(size_t)(c0 + c) * D + row0 + lr] = acc;
        }
      }
    }
    if (t < ntiles) {
      const int c0 = t * p.stage_rows, ct = min(p.stage_rows, p.C - c0);
      const int ncb = (ct + kCT - 1) / kCT;
      const float* G = s_G + (size_t)(t & 1) * kStageFloats;
      float* red = s_red + (size_t)(t & 1) * kMaxUnits * 16;
      auto run = [&](int gc, int cb) {
        load_block(gc);
        const int g = gc / nchunks, ch = gc - g * nchunks;
        float* out = red + ((size_t)(cb * ngroups + g) * nchunks + ch) * 16;
        if (p.nq == 4) project_unit<4>(a, G, D, ch * chunk, cb * kCT, ct, lane, out);
        else if (p.nq == 2) project_unit<2>(a, G, D, ch * chunk, cb * kCT, ct, lane, out);
        else project_unit<1>(a, G, D, ch * chunk, cb * kCT, ct, lane, out);
      };
      if (GC >= kWarps) {
        for (int gc = warp; gc < GC; gc += kWarps)
          for (int cb = 0; cb < ncb; ++cb) run(gc, cb);
      } else if (warp / GC < wpg) {
        for (int cb = warp / GC; cb < ncb; cb += wpg) run(warp % GC, cb);
      }
    }
    if (stamp && t < 2) p.ws_ts[11 + 2 * t] = tc::globaltimer_ns();
  }
  if (stamp) p.ws_ts[7] = tc::globaltimer_ns();
  if (tid == 0) p.ws_ts[336 + blockIdx.x] = tc::globaltimer_ns();     // every CTA's end
}

struct GsPlan {
  int grid, rows_per_cta, nb, rows_per_nb, stage_rows, nq;
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
  if (D > kMaxD) return MLA_E_SHAPE;
  size_t smem = ((size_t)2 * D + (size_t)rpc * D + (size_t)2 * kStageFloats + (size_t)2 * kMaxUnits * 16 + 32) * sizeof(float);
  if (smem + 256 > (size_t)di.smem_optin) return MLA_E_SHAPE;
  // projection geometry: column chunks of nq * 128, gradient rows staged per tile (a multiple of kCT that fits one staging
  // buffer and keeps a tile's (row group, chunk, class block) units within the partial-sum scratch)
  const int nq = D > 1024 ? 4 : (D > 512 ? 2 : 1);
  const int ngroups = (rpc + kRB - 1) / kRB, nchunks = (D + nq * 128 - 1) / (nq * 128);
  int stage_rows = std::min(kMaxStageRows, kStageFloats / D) / kCT * kCT;
  while (stage_rows > kCT && ngroups * nchunks * (stage_rows / kCT) > kMaxUnits) stage_rows -= kCT;
  if (stage_rows < kCT || ngroups * nchunks * (stage_rows / kCT) > kMaxUnits) return MLA_E_SHAPE;
  // raw-feature reduction: CTA = (batch slice, 256-column chunk). The number of slices nb (= partial rows every CTA sums
  // after the first barrier) is bounded by the CTAs available per chunk, by >= 8 rows per slice and by 160 KB of partials
  const int cch = (D + 255) / 256;
  int nb_cap = std::min(std::max(1, grid / cch), std::max(1, B / 8));
  nb_cap = std::max(1, std::min(nb_cap, (160 * 1024) / (D * 4)));
  if (cch > grid) return MLA_E_SHAPE;
  int rows_per_nb = (B + nb_cap - 1) / nb_cap;
  int nb = (B + rows_per_nb - 1) / rows_per_nb;
  pl->grid = grid; pl->rows_per_cta = rpc; pl->nb = nb; pl->rows_per_nb = rows_per_nb; pl->smem = smem;
  pl->stage_rows = stage_rows; pl->nq = nq;
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
  prm.stage_rows = pl.stage_rows; prm.nq = pl.nq;
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
