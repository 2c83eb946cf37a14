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
 * Supported: D % 4 == 0, 1 <= C <= 1024, label in [0, C). Two launches, deterministic.
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
MLA_API int    mla_conv2d_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, int accumulate, void* stream);
MLA_API size_t mla_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad);
MLA_API int    mla_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLA_B200_H */
