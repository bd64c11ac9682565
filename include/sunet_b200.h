/* sunet_b200.h - C ABI of the B200-native SUNet forward path.
 *
 * The reference (mehrdad78/SUNet_TF) is pure PyTorch and has no FFI; the "plugin API" of its forward path is the
 * nn.Module surface of model/SUNet_detail.py.  Each entry point below replaces the forward of one of those modules
 * (file:line cited per function).  sunet_tf_b200/modules.py binds them with ctypes and re-exposes the reference's
 * class names / constructor signatures / state_dict keys.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*.
 *   - return value: 0 OK, <0 argument/shape/alignment error, >0 cudaError_t.  sunet_last_error() returns the
 *     thread-local message.  Nothing throws across the boundary.
 *   - tensors are fp32, contiguous, resident on the current CUDA device, laid out exactly as the reference module
 *     receives / returns them.  All calls are asynchronous on `stream`.
 *   - parameters are handed over ONCE at pre-pack time by state_dict key (un-folded, as the reference stores them);
 *     the library keeps its own fp16 / folded device copies.  Handles are immutable afterwards and may be used from
 *     one host thread per GPU.
 *   - whole-model calls never allocate: the caller passes a workspace of sunet_workspace_bytes().  Per-module calls
 *     take scratch from the stream-ordered allocator (cudaMallocAsync) - they exist for drop-in use and parity
 *     tests at module granularity, not for throughput.
 */
#ifndef SUNET_B200_H
#define SUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sunet_handle_s* sunet_handle_t;

int sunet_abi_version(void);
const char* sunet_last_error(void);

/* Pre-pack the parameters of one module.  `kind` selects the module, `iargs`/`fargs` its constructor arguments,
 * (names[i], ptrs[i], numels[i]) its fp32 device parameters by state_dict key (relative to the module).
 *
 *  kind                iargs                                   fargs        replaces (model/SUNet_detail.py)
 *  "swin_block"        dim, H, W, num_heads, shift_size        qk_scale     SwinTransformerBlock :157-264
 *  "window_attention"  dim, num_heads                          qk_scale     WindowAttention      :59-138
 *  "mlp"               in, hidden, out                         -            Mlp                  :8-24
 *  "patch_merging"     dim, H, W                               -            PatchMerging         :285-322
 *  "upsample"          in_channels, factor(2|4), H, W          -            UpSample             :335-386
 *  "patch_embed"       in_chans, embed_dim, patch, has_norm    -            PatchEmbed           :518-556
 *  "sunet"             img_size, patch, in_chans, out_chans,   qk_scale     SUNet                :566-755
 *                      embed_dim, window, depth0..3, heads0..3
 * qk_scale <= 0 means head_dim ** -0.5 (the reference's `qk_scale or head_dim ** -0.5`, :80).
 */
int sunet_prepack(const char* kind, const int64_t* iargs, int n_iargs, const double* fargs, int n_fargs,
                  const char* const* names, const void* const* ptrs, const int64_t* numels, int n_params, void* stream,
                  sunet_handle_t* out);
int sunet_destroy(sunet_handle_t h);

/* SwinTransformerBlock.forward (:227-264): x (B, H*W, C) -> out (B, H*W, C). */
int sunet_swin_block_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream);
/* The same block on the library's own activation format - fp16 token rows (B*H*W, C), image order - without the fp32 casts and
 * the stream-ordered allocation of the entry point above: the form the whole-model forward runs, exposed for measurement
 * (BASELINE config 4) and for callers that keep an fp16 stream.  part 0: the whole block (:227-264), out may alias x;
 * part 1: only its window-attention part up to the per-head attention output (norm1, cyclic shift, window partition, qkv,
 * QK^T + relative-position bias + shifted-window mask, softmax, AV, window reverse, un-shift; :233-257 and :107-135, i.e.
 * WindowAttention.forward without its proj Linear, which the fused design runs inside the MLP kernel), out must not alias x.
 * The workspace (sunet_swin_block_f16_workspace_bytes) is caller-owned; nothing is allocated on the call. */
size_t sunet_swin_block_f16_workspace_bytes(sunet_handle_t h, int batch);
int sunet_swin_block_f16(sunet_handle_t h, const void* x, int batch, int part, void* out, void* workspace, size_t workspace_bytes,
                         void* stream);
/* WindowAttention.forward (:107-138): x (num_windows_total, 64, C); mask (mask_nw, 64, 64) fp32 or NULL. */
int sunet_window_attention_fwd(sunet_handle_t h, const float* x, int64_t num_windows, const float* mask, int mask_nw,
                               float* out, void* stream);
/* Mlp.forward (:18-24): x (rows, in) -> out (rows, out). */
int sunet_mlp_fwd(sunet_handle_t h, const float* x, int64_t rows, float* out, void* stream);
/* PatchMerging.forward (:301-322): x (B, H*W, C) -> out (B, H*W/4, 2C). */
int sunet_patch_merging_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream);
/* UpSample.forward (:365-386): x (B, H*W, C) -> factor 2: (B, 4*H*W, C/2); factor 4: (B, 4H, 4W, C). */
int sunet_upsample_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream);
/* PatchEmbed.forward (:548-556): x NCHW (B, in_chans, Himg, Wimg) -> out (B, Himg/p*Wimg/p, embed_dim). */
int sunet_patch_embed_fwd(sunet_handle_t h, const float* x, int batch, int himg, int wimg, float* out, void* stream);

/* SUNet.forward (:748-755) incl. the 1->3 channel repeat of SUNet_model.forward (model/SUNet.py:26-30):
 * x NCHW (B, in_chans in {1,3}, img, img) fp32 -> out NCHW (B, out_chans, img, img) fp32.
 * Batches larger than `max_chunk` images are processed in chunks of max_chunk inside the call (same workspace). */
size_t sunet_workspace_bytes(sunet_handle_t h, int batch, int max_chunk);
int sunet_forward(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, float* out, void* workspace,
                  size_t workspace_bytes, void* stream);
/* The demo.py I/O edge (demo.py:70-79: PIL RGB -> TF.to_tensor -> model -> torch.clamp(0,1) -> img_as_ubyte) in one call:
 * x (B, img, img, in_chans) uint8 interleaved as PIL / cv2 hold it -> out (B, img, img, out_chans) uint8 =
 * rint(clamp(model(x / 255), 0, 1) * 255).  The /255 is applied as the first kernel loads the image, clamp + quantisation
 * as the last kernel stores it; host<->device traffic is 1 byte per sample instead of 4. */
int sunet_forward_u8(sunet_handle_t h, const uint8_t* x, int in_chans, int batch, int max_chunk, uint8_t* out, void* workspace,
                     size_t workspace_bytes, void* stream);
/* The validation-loop forward (train.py:432-443): logits = model(x); prob = sigmoid(logits); se = (logits - target)^2,
 * reduced inside the last kernel.  target (B, target_chans, img, img) fp32 with target_chans == out_chans, or 3 with
 * out_chans == 1 (reduced to luminance 0.2989 R + 0.5870 G + 0.1140 B, :437-438); weight (B, 1, img, img) fp32 or NULL
 * (the per-pixel map of make_weights_from_numpy, :226-249, computed by the caller as in the reference); prob may be NULL.
 * sums (device, 5 doubles, overwritten): sum se, sum se*w, sum w, sum sqrt((logits-target)^2 + eps^2)*w, element count, so that
 * se.mean() = s0/s4 (:442), weighted MSE = s1/max(1e-8, s2) (:445), charbonnier_loss = s3/max(1e-8, s2) (:187-192, :447). */
int sunet_forward_eval(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, const float* target, int target_chans,
                       const float* weight, float eps, float* logits, float* prob, double* sums, void* workspace,
                       size_t workspace_bytes, void* stream);
/* Same forward with every kernel launch bracketed by CUDA events on `stream` (synchronises before returning).
 * recs[i] = {kind, device ms, algorithmic FLOPs, algorithmic bytes} in launch order; kind: 0 tcgen05 GEMM, 1 attention core,
 * 2 LayerNorm, 3 merge-gather+LN, 4 patch-embed conv, 5 up-sample combine, 6 tail stencil, 7 cast, 8 im2col, 9 fused LN+MLP+residual,
 * 10 fused LN+shift/partition+QKV+window attention, 11 fused x4 pixel-shuffle branch + folded output taps. */
typedef struct sunet_prof_rec { int kind; float ms; double flops; double bytes; } sunet_prof_rec;
int sunet_forward_profile(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, float* out, void* workspace,
                          size_t workspace_bytes, void* stream, sunet_prof_rec* recs, int max_recs, int* n_recs);
/* number of kernels one sunet_forward of `batch` images launches (for bench.py's gpu_launches) */
int64_t sunet_forward_launches(sunet_handle_t h, int batch, int max_chunk);

/* Any-resolution tile pipeline (demo_any_resolution.py:35-52, :125-139), device side.
 * sunet_tiles_extract: img NCHW (1, C, h, w) -> tiles (n*n, C, k, k) cut from the zero-padded centred canvas of side X.
 * sunet_tiles_fold:    tiles (count, C, k, k) covering tile indices [first, first+count) are accumulated into
 *                      acc (C, X, X) fp32 (must be zeroed by the caller before the first call);
 * sunet_tiles_finish:  out (1, C, h, w) = clamp(acc / cover_count, 0, 1) cropped to the image region. */
int sunet_tiles_extract(const float* img, int chans, int h, int w, int kernel, int stride, float* tiles, int first, int count,
                        void* stream);
int sunet_tiles_fold(const float* tiles, int chans, int h, int w, int kernel, int stride, int first, int count, float* acc,
                     void* stream);
int sunet_tiles_finish(const float* acc, int chans, int h, int w, int kernel, int stride, float* out, void* stream);

/* bring-up / microbenchmarks */
int sunet_selftest_umma(void* stream);
/* C[M,N] (fp16) = A[M,K] (fp16) * W[N,K]^T (fp16) + bias; used by tests to pin the tcgen05 GEMM in isolation */
int sunet_gemm_f16(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K, int act, void* stream);
/* out[rows,C] (fp16) = x + fc2(GELU(fc1(LayerNorm(x)))) on fp16 rows with fp32 device parameters (norm2 + Mlp + residual of
 * SwinTransformerBlock.forward :262) - the fused tcgen05 kernel in isolation (C in {96, 192}); packs, runs, frees. */
int sunet_ln_mlp_residual_f16(const void* x, int64_t rows, int C, const float* gamma, const float* beta, const float* w1,
                              const float* b1, const float* w2, const float* b2, void* out, void* stream);

/* The two remaining pieces of SUNet.forward_up_features on fp16 rows, in isolation (parity tests at module granularity):
 * out[rows,C] = LayerNorm(x) - the model's `norm` (:677, :718) and `norm_up` (:678, :732);
 * out[rows,C] = Linear(2C -> C)(cat([x, skip], -1)) - `concat_back_dim[i]` (:652-654, :728-729) as one tcgen05 GEMM with two
 * K segments (w fp32 [C][2C], b fp32 [C] or NULL; packs, runs, synchronises, frees). */
int sunet_layernorm_f16(const void* x, int64_t rows, int C, const float* gamma, const float* beta, void* out, void* stream);
int sunet_concat_linear_f16(const void* x, const void* skip, int64_t rows, int C, const float* w, const float* b, void* out,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUNET_B200_H */
