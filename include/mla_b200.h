/*
 * mla_b200.h — C ABI of libmla_b200.so: the sm_100a kernels behind MLA's alternating
 * unimodal training step (--gs_flag) and its test-time fusion (--dynamic).
 *
 * The reference (Cecile-hi/Multimodal-Learning-with-Alternating-Unimodal-Adaptation) is
 * pure Python and has no FFI of its own; each entry point below replaces the body of one
 * reference function (cited as file:line relative to the reference root) and is what a
 * ctypes stub on the reference side binds.
 *
 * Conventions
 *  - every function returns int: 0 = ok, <0 = argument error (MLA_E_*), >0 = cudaError_t.
 *    No C++ exception crosses the boundary.
 *  - the caller owns every buffer (device pointers unless stated); the library never
 *    allocates or frees device memory; scratch comes in through (ws, ws_bytes), sized by
 *    the matching *_workspace_bytes() query.
 *  - all work is asynchronous on the caller's stream (a cudaStream_t passed as void*);
 *    no hidden synchronisation.
 *  - row-major fp32 tensors, 16-byte aligned base pointers, labels int64.
 *  - the library holds no mutable global state except cached device attributes and
 *    per-kernel function attributes; it is re-entrant across distinct buffers/streams.
 */
#ifndef MLA_B200_H
#define MLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLA_ABI_VERSION 1

#if defined(__GNUC__)
#define MLA_API __attribute__((visibility("default")))
#else
#define MLA_API
#endif

#define MLA_E_BADARG   (-1)   /* null pointer / non-positive size / misaligned pointer   */
#define MLA_E_SHAPE    (-2)   /* shape outside what the kernel supports (see each call) */
#define MLA_E_WORKSPACE (-3)  /* ws == NULL or ws_bytes too small                        */
#define MLA_E_NODEVICE (-4)   /* no sm_100 device / cooperative launch unsupported      */

MLA_API int         mla_abi_version(void);
MLA_API const char* mla_error_string(int code);
/* Number of SMs on the current device (cached), or <0 on error. */
MLA_API int         mla_device_sm_count(void);
/* Kernel launches issued by this library since load (all entry points), for bench.py. */
MLA_API uint64_t    mla_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * GSPlugin.before_update body — utils/utils.py:34-41.
 *   r  = mean_b(feat)                       utils.py:34   (or feat_sum * inv_batch)
 *   k  = P r^T                              utils.py:35
 *   P' = P - (k k^T) ./ (alpha + k r)       utils.py:36   mode 0: ELEMENTWISE denominator
 *        P - (k k^T) /  (alpha + r k)                     mode 1: scalar OWM denominator
 *   P  = P' / ||P'||_F                      utils.py:38-40
 *   grad_w = grad_w @ P^T                   utils.py:41
 * The name gate (utils.py:32) and the counter gate (utils.py:29) stay on the host side.
 * Exactly one of feat (B x D) / feat_sum (D, the already reduced sum over the GLOBAL
 * batch) is non-NULL. inv_batch = 1 / B_global. P (D x D) and grad_w (C x D) are updated
 * in place. grad_w may be NULL (C ignored): P is updated, nothing is projected.
 * Supported: D % 4 == 0, 4 <= D <= 2048 with 148 SMs (the P slice of a CTA must fit in
 * shared memory), 1 <= C <= 4096, B >= 1. One cooperative launch, deterministic.
 */
MLA_API size_t mla_gs_project_workspace_bytes(int B, int D, int C);
MLA_API int    mla_gs_project(float* P, const float* feat, const float* feat_sum, float inv_batch,
                      float alpha, float* grad_w, int B, int D, int C, int mode,
                      void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * One modality turn of the shared head — main.py:432-435 (fc_out + CrossEntropyLoss +
 * backward restricted to the head), models/fusion_modules.py:19.
 *   logits = feat W^T + b ; loss = mean_b CE(logits, label)
 *   dlogits = (softmax(logits) - onehot) * grad_scale        (grad_scale = 1 / B_global)
 *   dW = dlogits^T feat ; db = sum_b dlogits ; dfeat = dlogits W ; feat_sum = sum_b feat
 * loss is the LOCAL mean (sum_b / B). Any of dW, db, dfeat, feat_sum, logits may be NULL
 * to skip that output (forward-only: pass dW = db = dfeat = NULL).
 * Supported: D % 4 == 0, 1 <= C <= 1024, label in [0, C). Three (C <= 16) to six launches, deterministic.
 */
MLA_API size_t mla_head_ce_workspace_bytes(int B, int D, int C);
MLA_API int    mla_head_ce(const float* feat, const float* W, const float* bias, const int64_t* label,
                   int B, int D, int C, float* logits, float* loss, float* dW, float* db,
                   float* dfeat, float* feat_sum, float grad_scale,
                   void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Test-time fusion + accuracy counters — main.py:65-106 (calculate_entropy,
 * calculate_gating_weights[3]) and main.py:640-676.
 *   dynamic != 0: H_m = -sum_{b,c} p log p, p = softmax(logits_m, dim=0) (over the BATCH,
 *                 as the reference does); w = softmax(-H) over modalities
 *   dynamic == 0: w = fixed_w (M host floats; main.py:648-651)
 *   fused = sum_m w_m logits_m ; argmax[0] = argmax_c fused, argmax[1+m] = argmax_c logits_m
 *   num[label] += 1 ; hits[j][label] += (argmax[j] == label)     (accumulated, not reset)
 * logits: HOST array of M device pointers (B x C each), 1 <= M <= 4. fused (B x C),
 * w_out (M), entropy_out (M, the scalar H_m of calculate_entropy, main.py:65-70; written
 * only when dynamic), argmax ((M+1) x B int32), hits ((M+1) x C int64), num (C int64); label,
 * hits and num may be NULL together (no counting). NaN weights propagate as in the
 * reference (0 * log 0). One cooperative launch, deterministic.
 */
MLA_API size_t mla_fuse_eval_workspace_bytes(int M, int B, int C);
MLA_API int    mla_fuse_eval(const float* const* logits, int M, int B, int C, int dynamic,
                     const float* fixed_w, const int64_t* label, float* fused, float* w_out,
                     float* entropy_out, int32_t* argmax, int64_t* hits, int64_t* num,
                     void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * ResNet-18 encoder convolutions — models/backbone.py:39-50 (BasicBlock 3x3 convs), :126-129
 * (1x1/2 downsample) and their backward (what loss.backward() runs through cuDNN in the
 * reference, main.py:435). Implicit GEMM on tcgen05 (TF32 operands, fp32 accumulate in TMEM).
 * Activations NHWC fp32; weights [Cout][R][S][Cin] fp32 (= torch channels_last memory of the
 * reference's OIHW parameter). No bias (the reference's convs have none).
 * Supported: R == S in {1, 3}, stride in {1, 2}, pad <= R/2, Cin and Cout in {64, 128k}.
 *   fprop : y  [N,OH,OW,Cout] = conv(x [N,H,W,Cin], w)
 *   dgrad : dx [N,H,W,Cin]    (+)= conv_transpose(dy [N,OH,OW,Cout], w)   (accumulate != 0: +=)
 *   wgrad : dw [Cout,R,S,Cin] = sum over pixels; split-K partials go to ws and are reduced in
 *           a fixed order (deterministic).
 */
MLA_API int    mla_conv2d_fprop(const float* x, const float* w, float* y, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, void* stream);
/* fprop that also emits BatchNorm partial sums of its output from the fp32 accumulators: stat_part
 * [mla_conv2d_fprop_stat_tiles(...)][2][Cout] floats = (sum y, sum y^2) per 128-row output tile; feed them to
 * mla_bn_stats_from_partials (saves the statistics pass over y). */
MLA_API int    mla_conv2d_fprop_stat_tiles(int N, int H, int W, int R, int S, int stride, int pad);
/* The same count for mla_conv2d_fprop16, whose 64 -> 64 channel 3x3 / stride-1 layers run on the halo-strip kernel with
 * resident weights (tiles of whole image rows instead of 128 output pixels). */
MLA_API int    mla_conv2d_fprop16_stat_tiles(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad);
MLA_API int    mla_conv2d_fprop_bnstats(const float* x, const float* w, float* y, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, float* stat_part, void* stream);
MLA_API int    mla_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, int accumulate, void* stream);
/* 2-byte operand variants (tcgen05 kind::f16, fp32 accumulate; half the k-blocks, MMA instructions and operand bytes):
 *   fprop16: x16 [N,H,W,Cin] fp16, w16 [Cout,R,S,Cin] fp16 -> y fp32 (+ optional BN partial sums as fprop_bnstats).
 *            fp16 carries TF32's 10-bit mantissa: the products are those of the TF32 path. Cin % 64 == 0.
 *   dgrad16: dy16 [N,OH,OW,Cout] bf16, wt16 [Cin,R,S,Cout] bf16 (the filter TRANSPOSED, mla_filter_transpose16; the
 *            hardware rejects mixed fp16 x bf16 operands) -> dx fp32 (accumulate != 0: +=). Cout, Cin % 64 == 0. */
MLA_API int    mla_conv2d_fprop16(const void* x16, const void* w16, float* y, int N, int H, int W, int Cin, int Cout,
                        int R, int S, int stride, int pad, float* stat_part, void* stream);
MLA_API int    mla_conv2d_dgrad16(const void* dy16, const void* wt16, float* dx, int N, int H, int W, int Cin, int Cout,
                        int R, int S, int stride, int pad, int accumulate, void* stream);
/*   wgrad16: x16 [N,H,W,Cin] bf16, dy16 [N,OH,OW,Cout] bf16 -> dw fp32 [Cout,R,S,Cin]; Cin, Cout % 64 == 0. */
MLA_API size_t mla_conv2d_wgrad16_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                        int pad);
MLA_API int    mla_conv2d_wgrad16(const void* x16, const void* dy16, float* dw, int N, int H, int W, int Cin, int Cout,
                        int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream);
/* dst16[i] = fp16(src[i]) (bf16 != 0: bfloat16), n % 4 == 0. */
MLA_API int    mla_cast16(const float* src, void* dst16, long long n, int bf16, void* stream);
/* wt16 [Cin,RS,Cout] fp16 (bf16 != 0: bfloat16) = transpose of w [Cout,RS,Cin] fp32. */
MLA_API int    mla_filter_transpose16(const float* w, void* wt16, int Cout, int RS, int Cin, int bf16, void* stream);
MLA_API size_t mla_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad);
MLA_API int    mla_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * m3ae transformer encoders (reference models/m3ae.py:86-179). The Linears run on the
 * convolution GEMMs above as 1x1 convolutions over an N=1 image of M x 1 pixels.
 *
 * Fused attention core (m3ae.py:103-121): out = softmax(scale * q k^T, padded keys FILLED with -1e7) v, per head.
 *   qkv [B, S, 3, H, Dh] fp32 (the qkv_linear output), key_mask [B, S] fp32 (> 0: padded key) or NULL,
 *   out [B, S, H*Dh] fp32, stats [B, H, S, 2] fp32 (row max, row sum of exp), qkv16 [B, S, 3, H, Dh] fp16 (written by
 *   the forward; stats and qkv16 are what the backward needs). Dh = 32 or 64. fp16 operands (10-bit mantissa, like
 *   TF32), fp32 accumulate and softmax; the S x S matrix is never written.
 * backward: dqkv [B, S, 3, H, Dh] fp32 (every element written); dout is rescaled by a power of two internally so that
 *   fp16 holds it; ws from mla_attention_backward_workspace_bytes. Deterministic (no atomics on the results). */
MLA_API int    mla_attention_forward(const float* qkv, const float* key_mask, float* out, float* stats, void* qkv16,
                        int B, int S, int H, int Dh, float scale, void* stream);
MLA_API size_t mla_attention_backward_workspace_bytes(int B, int S, int H, int Dh);
MLA_API int    mla_attention_backward(const void* qkv16, const float* key_mask, const float* out, const float* dout,
                        const float* stats, float* dqkv, int B, int S, int H, int Dh, float scale, void* ws,
                        size_t ws_bytes, void* stream);

/* Linear layer on the tcgen05 GEMM: y [M, N] = x16 [M, K] (fp16) * w16 [N, K]^T (fp16) + bias [N] + resid [M, N]
 * (bias / resid may be NULL; the adds happen in the GEMM epilogue). K, N % 64 == 0. nn.Linear of m3ae.py:72-73,98-99. */
MLA_API int    mla_linear_forward16(const void* x16, const void* w16, const float* bias, const float* resid, float* y,
                        int M, int K, int N, void* stream);
/* dx [M, K] = dy [M, N] * w [N, K], both fp32 already rounded to TF32 (mla_round_colsum / mla_round_tf32): the data
 * gradient of the same Linear. K, N % 64 == 0. (The weight gradient is mla_conv2d_wgrad with N=1, H=M, W=1, R=S=1.)
 * The two Linear entry points use tcgen05 CTA pairs (256-row MMAs); they must not run concurrently with OTHER pair
 * kernels on a second stream (MLA_LINEAR_PAIR=0 selects single-CTA tiles). */
MLA_API int    mla_linear_dgrad(const float* dy, const float* w, float* dx, int M, int K, int N, void* stream);
/* fp16 backward of a Linear (default of the transformer paths): the gradient operand is cast to fp16 after multiplying by a
 * power of two F chosen from its largest element (exact; fp16 then carries TF32's 10-bit mantissa at twice TF32's
 * tensor-core rate), and the GEMM epilogues multiply by 1/F.
 *   mla_grad_operand16: out16 [M, N] = fp16(F * v), v = dy (u == NULL) or dy * gelu'(u); colsum [N] = column sums of v (the
 *     bias gradient); scale_io = 2 floats, [1] receives 1/F. ws from mla_round_colsum_workspace_bytes(M, N).
 *   mla_linear_dgrad16: dx [M, K] = *out_scale * dy16 [M, N] * wt16 [K, N]^T, wt16 = mla_filter_transpose16(w, ., N, 1, K, 0).
 *   mla_linear_wgrad16: dw [N, K] = *out_scale * dy16^T x16, x16 [M, K] fp16; ws from
 *     mla_conv2d_wgrad16_workspace_bytes(1, M, 1, K, N, 1, 1, 1, 0). */
MLA_API int    mla_grad_operand16(const float* dy, const float* u, void* out16, float* colsum, float* scale_io, long long M,
                        int N, void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_linear_dgrad16(const void* dy16, const void* wt16, const float* out_scale, float* dx, int M, int K, int N,
                        void* stream);
MLA_API int    mla_linear_wgrad16(const void* x16, const void* dy16, const float* out_scale, float* dw, int M, int K, int N,
                        void* ws, size_t ws_bytes, void* stream);
/* nn.LayerNorm over the last dimension D (D % 4 == 0, D <= 1280), biased variance, as torch. Outputs (each may be NULL):
 * y fp32, y16 fp16 (the operand of the next GEMM), y_r fp32 rounded to TF32 (the operand of that GEMM's weight
 * gradient); mean / rstd [M] are kept for the backward pass. */
MLA_API int    mla_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps, long long M, int D,
                        float* y, void* y16, float* y_r, float* mean, float* rstd, void* stream);
/* dx = resid + dLN/dx (resid may be NULL: the residual branch's gradient), dgamma, dbeta [D]. Deterministic. */
MLA_API size_t mla_layernorm_backward_workspace_bytes(long long M, int D);
MLA_API int    mla_layernorm_backward(const float* dy, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* resid, long long M, int D, float* dx, float* dgamma,
                        float* dbeta, void* ws, size_t ws_bytes, void* stream);
/* x16 = fp16(f(x)), x_r = tf32(f(x)); f = identity, or exact (erf) GELU when gelu != 0. n % 4 == 0. */
MLA_API int    mla_cast_round(const float* x, void* x16, float* x_r, long long n, int gelu, void* stream);
/* out_r [M, N] = tf32(dy) (u == NULL) or tf32(dy * gelu'(u)) — the dy operand of a Linear's gradient GEMMs — and
 * colsum [N] = its column sums (the bias gradient), taken before rounding. Deterministic. */
MLA_API size_t mla_round_colsum_workspace_bytes(long long M, int N);
MLA_API int    mla_round_colsum(const float* dy, const float* u, float* out_r, float* colsum, long long M, int N, void* ws,
                        size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Memory-bound encoder pieces (NHWC fp32) — models/backbone.py:142-160 (ResNet.forward),
 * :36-52 (BasicBlock.forward), nn.BatchNorm2d / nn.MaxPool2d semantics, basic_model.py:56-65
 * (global average pool), and their backward passes.
 *
 * stem_im2col : 7x7/2 stem (backbone.py:78-83,149) as a GEMM: rows m = (n,oh,ow), columns
 *               k = (r*S+s)*Cin+ci, zero-padded to Kp. Reads the RAW input: element (n,ci,h,w) at
 *               in[(n/T)*sB + (n%T)*sT + ci*sC + h*W + w]  (frame fold backbone.py:144-147).
 * pad_rows    : [rows][k] <-> [rows][kp] zero-padded copy (stem weights / weight gradient).
 * bn_train_stats: batch mean / biased variance per channel over M rows; running stats updated
 *               in place with momentum and the UNBIASED variance; emits mean, invstd and the
 *               fused affine scale = gamma*invstd, shift = beta - mean*scale.
 * bn_eval_coeffs: scale/shift from the running statistics (eval mode).
 * bn_apply    : out = relu?( y*scale + shift + (res ? res*res_scale + res_shift : 0) );
 *               res_scale/res_shift NULL = plain identity shortcut.
 * bn_backward : g = dz * (z > 0) (z NULL: no mask); dgamma = sum g*xhat; dbeta = sum g;
 *               dy = gamma*invstd*(g - dbeta/M - xhat*dgamma/M); g_out (optional) receives g.
 * bn_relu_maxpool / maxpool_relu_backward: stem BN+ReLU+MaxPool(3,2,1) fused; idx holds the
 *               argmax window position (uint8 per element).
 * avgpool_*   : feat[b] = mean of `rows` consecutive NHWC rows (h*w, or T*h*w for video).
 * ws for the BN calls: mla_bn_workspace_bytes(M, C) bytes, ZERO-FILLED by the caller before its first use
 *               (it holds the completion tickets of the single-launch reductions; every call leaves
 *               them at zero again, so one buffer serves any sequence of calls on one stream).
 *               C % 64 == 0. Deterministic: partial sums are combined in block order.
 */
MLA_API int    mla_stem_im2col(const float* in, float* col, int N, int T, long long sB, long long sT,
                        long long sC, int Cin, int H, int W, int R, int S, int stride, int pad, int Kp,
                        void* stream);
/* The same matrix as 2-byte copies: col16 fp16 (the fprop16 operand) and col16b bf16 (the wgrad16 operand; may be NULL,
 * e.g. in evaluation). Kp % 64 == 0. */
MLA_API int    mla_stem_im2col16(const float* in, void* col16, void* col16b, int N, int T, long long sB, long long sT,
                        long long sC, int Cin, int H, int W, int R, int S, int stride, int pad, int Kp, void* stream);
/* dst = round-to-nearest-TF32(src), n % 4 == 0 (weights before they feed the tensor cores). */
MLA_API int    mla_round_tf32(const float* src, float* dst, long long n, void* stream);
MLA_API int    mla_pad_rows(const float* src, float* dst, int rows, int k, int kp, int unpad, void* stream);
MLA_API size_t mla_bn_workspace_bytes(long long M, int C);
MLA_API int    mla_bn_train_stats(const float* y, long long M, int C, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float momentum, float eps,
                        float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                        void* ws, size_t ws_bytes, void* stream);
/* Same outputs as mla_bn_train_stats from the per-tile partial sums of mla_conv2d_fprop_bnstats (M = number of
 * pixels the statistics cover; ws: mla_bn_workspace_bytes(ntiles, C), zero-filled like the other BN workspaces). */
MLA_API int    mla_bn_stats_from_partials(const float* part, int ntiles, long long M, int C, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                        float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                        void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, int C, float* scale_out, float* shift_out,
                        void* stream);
MLA_API int    mla_bn_apply(const float* y, const float* scale, const float* shift, const float* res,
                        const float* res_scale, const float* res_shift, int relu, float* out,
                        long long M, int C, void* stream);
/* Same as mla_bn_apply, additionally writing the sign of the output as a bitmask (1 bit per element, element e =
 * bit e % 32 of word e / 32; M*C/32 words, C % 32 == 0): mla_bn_backward_mask reads it in place of the 32x larger
 * activation z when it masks the gradient of the ReLU. */
MLA_API int    mla_bn_apply_mask(const float* y, const float* scale, const float* shift, const float* res,
                        const float* res_scale, const float* res_shift, int relu, float* out,
                        unsigned int* relu_mask, long long M, int C, void* stream);
MLA_API int    mla_bn_backward_mask(const float* dz, const unsigned int* relu_mask, const float* y, const float* mean,
                        const float* invstd, const float* gamma, long long M, int C, float* dgamma,
                        float* dbeta, float* dy, float* g_out, void* ws, size_t ws_bytes, void* stream);
/* The same three passes with 2-byte side outputs for the kind::f16 convolutions. Any of relu_mask / out16 / out16b /
 * dy16 / z may be NULL, and so may the fp32 out / dy when a 2-byte output is given: out16 = fp16 copy of the activation
 * (fprop16 operand), out16b = bf16 copy (wgrad16 x operand), dy16 = bf16 gradient (dgrad16 / wgrad16 operand). */
MLA_API int    mla_bn_apply_ex(const float* y, const float* scale, const float* shift, const float* res,
                        const float* res_scale, const float* res_shift, int relu, float* out,
                        unsigned int* relu_mask, void* out16, void* out16b, long long M, int C, void* stream);
MLA_API int    mla_bn_backward_ex(const float* dz, const float* z, const unsigned int* relu_mask, const float* y,
                        const float* mean, const float* invstd, const float* gamma, long long M, int C,
                        float* dgamma, float* dbeta, float* dy, void* dy16, float* g_out, void* ws,
                        size_t ws_bytes, void* stream);
MLA_API int    mla_bn_relu_maxpool_ex(const float* y, const float* scale, const float* shift, float* out, void* out16,
                        void* out16b, unsigned char* idx, int N, int H, int W, int C, void* stream);
MLA_API int    mla_bn_backward(const float* dz, const float* z, const float* y, const float* mean,
                        const float* invstd, const float* gamma, long long M, int C, float* dgamma,
                        float* dbeta, float* dy, float* g_out, void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_bn_relu_maxpool(const float* y, const float* scale, const float* shift, float* out,
                        unsigned char* idx, int N, int H, int W, int C, void* stream);
MLA_API int    mla_maxpool_relu_backward(const float* dp, const float* p, const unsigned char* idx, float* g,
                        int N, int H, int W, int C, void* stream);
MLA_API int    mla_avgpool_forward(const float* fm, float* feat, int B, int rows, int C, void* stream);
MLA_API int    mla_avgpool_backward(const float* dfeat, float* dfm, int B, int rows, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * fp16 gradient operands with an exact power-of-two scale (the backward convolutions of backbone.py:39-50 under autograd,
 * main.py:435): TF32's 10-bit operand mantissa at the kind::f16 tensor-core rate.
 *   mla_bn_backward_f16     BatchNorm backward whose dy leaves as fp16(dy * F); F = 2^k is chosen on the device from
 *                           max |dz * relu mask| * |gamma| * invstd (found by the reduction pass) so that the bound lands in
 *                           [2^8, 2^9); conversions saturate. gscale[0] = F, gscale[1] = 1 / F (device floats).
 *   mla_conv2d_dgrad16_f16  dx (+)= *out_scale * (dy16 (*) wt16)   dy16 fp16 scaled, wt16 fp16 transposed filter [Cin][R][S][Cout]
 *   mla_conv2d_wgrad16_f16  dw = *out_scale * (dy16^T (*) x16)     x16 = the forward's fp16 activation; workspace as
 *                           mla_conv2d_wgrad16_workspace_bytes
 *   mla_filter_transpose16_batch  every filter of an encoder transposed / cast in one launch; seg_table = device array of
 *                           {int64 element offset, int32 Cout, RS, Cin, first tile} (24 bytes each), ntiles = all 32x32 tiles
 */
MLA_API int    mla_bn_backward_f16(const float* dz, const unsigned int* relu_mask, const float* y, const float* mean,
                        const float* invstd, const float* gamma, long long M, int C, float* dgamma, float* dbeta,
                        void* dy16, float* g_out, float* gscale, void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_conv2d_dgrad16_f16(const void* dy16, const void* wt16, const float* out_scale, float* dx, int N, int H,
                        int W, int Cin, int Cout, int R, int S, int stride, int pad, int accumulate, void* stream);
MLA_API int    mla_conv2d_wgrad16_f16(const void* x16, const void* dy16, const float* out_scale, float* dw, int N, int H,
                        int W, int Cin, int Cout, int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_filter_transpose16_batch(const float* flat, void* flat_t16, const void* seg_table, int nseg, int ntiles,
                        int bf16, void* stream);

/* ---------------------------------------------------------------------------------------
 * The 7x7 / stride 2 / pad 3 stem convolution without an im2col matrix (stem_s2d.cu) — models/backbone.py:78-83,149,
 * 150-152 (conv1 -> bn1 -> relu -> maxpool) and their autograd. Cin <= 4, 64 output channels, fp16 operands.
 *   mla_stem_s2d_pack      raw NCHW / NCTHW fp32 input (strides as in mla_stem_im2col) -> the zero-padded space-to-depth
 *                          tensor xs16 [N][OH+3][OW+3][16] fp16 (mla_stem_s2d_input_elems elements): cell (r, q) channel
 *                          (pr*2+pc)*Cin + c = x[n][c][2r+pr-3][2q+pc-3]
 *   mla_stem_s2d_weights   w [64][7][7][Cin] fp32 (channels_last OIHW) -> w2_16 [64][4][4][16] fp16, regrouped the same way
 *   mla_stem_s2d_fprop     y16 [N][OH][OW][64] fp16 = conv(xs16, w2_16); stat_part (may be NULL) receives the BatchNorm
 *                          partial sums of the fp32 accumulators, [mla_stem_s2d_tiles][2][64] (-> mla_bn_stats_from_partials)
 *   mla_stem_s2d_wgrad     dw [64][7][7][Cin] = *out_scale * sum over pixels dy16^T patch(xs16); deterministic
 *   mla_bn_relu_maxpool16  mla_bn_relu_maxpool_ex over the fp16 y16; idx bit 4 marks pooled values that are not positive
 *   mla_pool_bn_backward_f16  BatchNorm backward of the stem with the ReLU + MaxPool backward folded in: dp = gradient of the
 *                          POOLED activation [N][PH][PW][C] fp32, idx from mla_bn_relu_maxpool16, (H, W) = pre-pool extent;
 *                          dy16 = fp16(dy * F), gscale = {F, 1 / F} as in mla_bn_backward_f16. ws: mla_bn_workspace_bytes.
 */
MLA_API long long mla_stem_s2d_input_elems(int N, int H, int W);
MLA_API int    mla_stem_s2d_tiles(int N, int H, int W);
MLA_API int    mla_stem_s2d_pack(const float* in, void* xs16, int N, int T, long long sB, long long sT, long long sC, int Cin,
                        int H, int W, void* stream);
MLA_API int    mla_stem_s2d_weights(const float* w, void* w2_16, int Cin, void* stream);
MLA_API int    mla_stem_s2d_fprop(const void* xs16, const void* w2_16, void* y16, int N, int H, int W, float* stat_part,
                        void* stream);
MLA_API size_t mla_stem_s2d_wgrad_workspace_bytes(void);
MLA_API int    mla_stem_s2d_wgrad(const void* xs16, const void* dy16, const float* out_scale, float* dw, int N, int H, int W,
                        int Cin, void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_bn_relu_maxpool16(const void* y16, const float* scale, const float* shift, float* out, void* out16,
                        unsigned char* idx, int N, int H, int W, int C, void* stream);
MLA_API int    mla_pool_bn_backward_f16(const float* dp, const unsigned char* idx, const void* y16, const float* mean,
                        const float* invstd, const float* gamma, int N, int H, int W, int C, float* dgamma, float* dbeta,
                        void* dy16, float* gscale, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * OGM / OGM-GE gradient modulation of the joint-training step — main.py:312-410.
 *   mla_ogm_scores   score[m] = sum_b softmax(logits[m])[b][label[b]], b added in index order   main.py:315-317, 373-374
 *                    `logits` is a HOST array of M (2 or 3) device pointers to B x C matrices; ws >= M * B floats.
 *   mla_ogm_coeff    coeff[m] from the (global-batch) scores: M = 2 -> main.py:376-384, M = 3 -> main.py:319-334;
 *                    score / coeff are device arrays, the branch runs on the device (no host read).
 *   mla_ogm_modulate grad[seg_off[i] .. + seg_len[i]) = grad * *coeff (+ noise * seg_std[i])    main.py:393-408
 *                    over `nseg` segments of one flat gradient buffer (the 4-D parameters of one encoder); seg_off /
 *                    seg_len are DEVICE arrays (elements), max_len = the longest segment; noise (same layout as grad,
 *                    N(0,1) draws) and seg_std (std(grad) + 1e-8 per segment) are both NULL for plain OGM.
 */
MLA_API size_t mla_ogm_scores_workspace_bytes(int M, int B);
MLA_API int    mla_ogm_scores(const float* const* logits, int M, const int64_t* label, int B, int C, float* score,
                        void* ws, size_t ws_bytes, void* stream);
MLA_API int    mla_ogm_coeff(const float* score, int M, float alpha, float* coeff, void* stream);
MLA_API int    mla_ogm_modulate(float* grad, const long long* seg_off, const long long* seg_len, int nseg,
                        long long max_len, const float* coeff, const float* noise, const float* seg_std, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dataset tuple producer, visual modality — replaces, per batch, the per-sample CPU pipeline of
 * dataset/dataset.py:123-161 (same Compose at :448-480, :753-803): PIL RandomResizedCrop(+ RandomHorizontalFlip) or
 * Resize((OH, OW)) -> ToTensor -> Normalize -> T frames stacked on dim 1.
 *   src        packed uint8 HWC RGB frames (device), src_bytes long
 *   desc       DEVICE array [nframes][14] int32: {src offset lo, hi (bytes), H, W, crop top, left, crop_h, crop_w, flip,
 *              destination slot b * T + t, RH, RW, oy, ox}: the crop box is resampled to RH x RW and the [OH, OW] output is
 *              the window of that image at (oy, ox). Resize((OH, OW)) / RandomResizedCrop: crop box (0, 0, H, W) or the
 *              drawn one, RH = OH, RW = OW, oy = ox = 0. Resize(s) + CenterCrop(s) (dataset.py:251-256, 414-421): shorter
 *              side -> s, (oy, ox) = the centre-crop origin.
 *   filter     0 = BILINEAR (AVDataset), 1 = BICUBIC (CAVDataset, M3AEDataset)
 *   mean3/std3 HOST arrays of 3 floats (read at launch)
 *   out        [B, 3, T, OH, OW] fp32 — bit-identical to torchvision's PIL path (Pillow's two-pass antialiased bilinear
 *              resampling with 8-bit intermediate and 22-bit fixed-point coefficients, float(u8) / 255, (x - mean) / std)
 *   status     optional DEVICE int: 0, or 1 + the index of an invalid descriptor (such frames are skipped)
 * crop / output ratios up to 15.5 (bilinear) or 7.5 (bicubic) per axis; larger ones are reported through status.
 */
MLA_API size_t mla_frames_to_batch_workspace_bytes(int nframes, int OH, int OW, int max_crop_h);
MLA_API int    mla_frames_to_batch(const unsigned char* src, long long src_bytes, const int* desc, int nframes, int B, int T,
                        int OH, int OW, int max_crop_h, int filter, const float* mean3, const float* std3, float* out,
                        int* status, void* ws, size_t ws_bytes, void* stream);

/* Audio member of the CAV-MAE-style tuples — dataset/dataset.py:281-294 (fbank_aug: torchaudio FrequencyMasking /
 * TimeMasking), :312-321 (normalisation, noise, roll) for a whole batch of pre-computed filterbank arrays [B][T][F]:
 *   out[b][(t + shift) mod T][f] = norm(masked(x[b][t][f])) + noise[b][t][f] * amp[b] / 10
 * params (DEVICE) [B][6] int32 = {f0, f1, t0, t1, shift, add_noise}: rows f0 <= f < f1 and frames t0 <= t < t1 are zeroed
 * BEFORE the normalisation (x - mean) / std (skipped when skip_norm); noise may be NULL. fp32, same roundings as torch. */
MLA_API int    mla_spec_to_batch(const float* fbank, const int* params, const float* amp, const float* noise, float mean,
                        float std, int skip_norm, int B, int T, int F, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLA_B200_H */
