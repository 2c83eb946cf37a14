// OGM / OGM-GE gradient modulation of the joint-training step — reference main.py:312-410:
//   score_m = sum_b softmax(out_m)[b][label_b]                       main.py:315-317 / 373-374
//   ratios, coefficient 1 - tanh(alpha * relu(ratio)) of the dominant modality     :319-334 / 376-384
//   conv (4-D) gradients of that modality's encoder *= coeff [+ N(0, std(grad) + 1e-8)]   :343-368 / 393-408
// The reference evaluates the scores with a per-sample Python loop (one device->host read per sample), branches on the
// host and rescales parameter by parameter; here: one kernel for the scores (fixed summation order b = 0, 1, ...,
// like the Python sum), one for the coefficients (device-side branch, no host read), one launch per encoder that
// rescales every selected segment of its flat gradient buffer.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int kScoreThreads = 256;
struct LogitPtrs { const float* p[3]; };     // passed by value: no device-side pointer table, graph-capturable

// grid = M (one block per modality). prob[b] = softmax(out[b])[label[b]] by a warp per sample, then thread 0 adds the B
// probabilities in index order (the reference's Python sum()).
__global__ void __launch_bounds__(kScoreThreads) ogm_scores_kernel(LogitPtrs logits, const int64_t* __restrict__ label,
                                                                   int B, int C, float* __restrict__ score,
                                                                   float* __restrict__ prob_ws) {
  const int m = blockIdx.x;
  const float* out = logits.p[m];
  float* prob = prob_ws + (size_t)m * B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kScoreThreads / 32;
  for (int b = warp; b < B; b += nw) {
    const float* row = out + (size_t)b * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
    mx = mla::warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - mx);
    s = mla::warp_sum(s);
    const long long lab = label[b];
    if (lane == 0) prob[b] = (lab >= 0 && lab < C) ? expf(row[lab] - mx) / s : __int_as_float(0x7fc00000);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += prob[b];
    score[m] = acc;
  }
}

// coeff[m] for M = 2 (a, v: main.py:376-384) or M = 3 (a, v, t: main.py:319-334). One thread.
__global__ void ogm_coeff_kernel(const float* __restrict__ score, int M, float alpha, float* __restrict__ coeff) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float sa = score[0], sv = score[1];
  if (M == 2) {
    const float ratio_v = sv / sa, ratio_a = 1.f / ratio_v;
    coeff[0] = coeff[1] = 1.f;
    if (ratio_v > 1.f) coeff[1] = 1.f - tanhf(alpha * fmaxf(ratio_v, 0.f));
    else coeff[0] = 1.f - tanhf(alpha * fmaxf(ratio_a, 0.f));
  } else {
    const float st = score[2];
    const float ratio_v = sv / (sa + st), ratio_a = sa / (sv + st), ratio_t = st / (sv + sa);
    coeff[0] = coeff[1] = coeff[2] = 1.f;
    if (ratio_v > 1.f) coeff[1] = 1.f - tanhf(alpha * fmaxf(ratio_v, 0.f));
    else if (ratio_t > 1.f) coeff[2] = 1.f - tanhf(alpha * fmaxf(ratio_t, 0.f));
    else coeff[0] = 1.f - tanhf(alpha * fmaxf(ratio_a, 0.f));
  }
}

// grad[seg] = grad[seg] * coeff (+ noise[seg] * seg_std[seg]); blockIdx.y = segment.
__global__ void __launch_bounds__(256) ogm_modulate_kernel(float* __restrict__ grad, const long long* __restrict__ seg_off,
                                                           const long long* __restrict__ seg_len,
                                                           const float* __restrict__ coeff, const float* __restrict__ noise,
                                                           const float* __restrict__ seg_std) {
  const int s = blockIdx.y;
  const long long off = seg_off[s], n = seg_len[s];
  const float c = __ldg(coeff);
  const float sd = (noise != nullptr) ? seg_std[s] : 0.f;
  float* g = grad + off;
  const float* z = (noise != nullptr) ? noise + off : nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = g[i] * c;                       // parms.grad * coeff      (main.py:398 / 401)
    if (z != nullptr) v += z[i] * sd;         // + normal_(0, std + 1e-8): z ~ N(0, 1) drawn by the caller's generator
    g[i] = v;
  }
}

}  // namespace

extern "C" size_t mla_ogm_scores_workspace_bytes(int M, int B) {
  if (M < 1 || B < 1) return 0;
  return (size_t)M * B * sizeof(float);
}

extern "C" int mla_ogm_scores(const float* const* logits, int M, const int64_t* label, int B, int C, float* score,
                              void* ws, size_t ws_bytes, void* stream) {
  if (!logits || !label || !score) return MLA_E_BADARG;
  if (M < 2 || M > 3 || B < 1 || C < 1) return MLA_E_SHAPE;
  const size_t need = mla_ogm_scores_workspace_bytes(M, B);
  if (!ws || ws_bytes < need) return MLA_E_WORKSPACE;
  for (int m = 0; m < M; ++m)
    if (!logits[m]) return MLA_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LogitPtrs h{};
  for (int m = 0; m < M; ++m) h.p[m] = logits[m];
  ogm_scores_kernel<<<M, kScoreThreads, 0, st>>>(h, label, B, C, score, static_cast<float*>(ws));
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

extern "C" int mla_ogm_coeff(const float* score, int M, float alpha, float* coeff, void* stream) {
  if (!score || !coeff) return MLA_E_BADARG;
  if (M < 2 || M > 3) return MLA_E_SHAPE;
  ogm_coeff_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(score, M, alpha, coeff);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}

extern "C" int mla_ogm_modulate(float* grad, const long long* seg_off, const long long* seg_len, int nseg,
                                long long max_len, const float* coeff, const float* noise, const float* seg_std,
                                void* stream) {
  if (!grad || !seg_off || !seg_len || !coeff || nseg < 1 || max_len < 1) return MLA_E_BADARG;
  if ((noise == nullptr) != (seg_std == nullptr)) return MLA_E_BADARG;
  const mla::DeviceInfo& di = mla::device_info();
  if (di.ok != 1) return di.ok;
  long long gx = (max_len + 256 * 8 - 1) / (256 * 8);
  const long long cap = std::max(1, 8 * di.sm_count / nseg);
  gx = std::max(1LL, std::min(gx, cap));
  ogm_modulate_kernel<<<dim3((unsigned)gx, (unsigned)nseg), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, seg_off, seg_len, coeff, noise, seg_std);
  MLA_CUDA_TRY(cudaGetLastError());
  mla::count_launch();
  return 0;
}
