// Bandwidth-bound kernels of the SUNet forward (everything that is not a GEMM or the attention core).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

int cast_f32_to_f16(const float* in, __half* out, int64_t n, cudaStream_t s);
int cast_f16_to_f32(const __half* in, float* out, int64_t n, cudaStream_t s);

// LayerNorm over the last dim (eps 1e-5, biased variance): nn.LayerNorm at SUNet_detail.py:192,198,677-678.
int layernorm_f16(const __half* in, int64_t ld_in, __half* out, int64_t ld_out, const float* gamma, const float* beta,
                  int64_t M, int C, cudaStream_t s);

// PatchMerging gather + LayerNorm(4C) (SUNet_detail.py:310-319): in (B,H,W,C) -> out (B*H/2*W/2, 4C), order TL,BL,TR,BR.
int merge_gather_ln_f16(const __half* in, __half* out, const float* gamma, const float* beta, int B, int H, int W, int C,
                        cudaStream_t s);

// conv_first (3x3, pad 1) folded with patch_embed.proj (4x4, stride 4) into one 6x6 / stride 4 / pad 1 conv, + LayerNorm.
// img: fp32 NCHW (B, img_chans in {1,3}, Himg, Wimg), or - img_fmt IMG_U8_NHWC - the 8-bit interleaved image as PIL holds it
// (B, Himg, Wimg, img_chans), scaled by 1/255 on load (TF.to_tensor, demo.py:71);
// wfold: fp32 [108][E] (tap-major: k = c*36 + u*6 + v); out fp16 (B, Himg/4*Wimg/4, E).
enum ImgFmt { IMG_F32_NCHW = 0, IMG_U8_NHWC = 1 };
// wpk (optional): the same weights as fp16 [E][PATCH_EMBED_WPK_PITCH] (pack_patch_embed_f16) for the tensor-core form of the kernel,
// taken when the token grid splits into 8 x 16 tiles (Wimg % 64 == 0); otherwise the fp32 FFMA form runs.
constexpr int PATCH_EMBED_WPK_PITCH = 120;
int patch_embed_fused(const void* img, int img_fmt, int img_chans, int B, int Himg, int Wimg, const float* wfold, const __half* wpk,
                      const float* bfold, const float* gamma, const float* beta, int E, __half* out, cudaStream_t s);
int pack_patch_embed_f16(const float* wfold, __half* wpk, int E, cudaStream_t s);

// im2col for a stand-alone PatchEmbed (conv k = stride = P): A[m][c*P*P + ky*P + kx] fp16
int im2col_patch(const float* img, int B, int Cin, int Himg, int Wimg, int P, __half* out, cudaStream_t s);

// Dual up-sample combine: out(b, r*h+i, r*w+j, :) = Yp[(b,h,w)][i*r+j][:] + bilinear_r(Z)(b, r*h+i, r*w+j, :)
// Yp: fp16 [B*H*W][r*r][Co] (pixel-shuffle branch after its folded 1x1 convs), Z: fp16 [B*H*W][Co] (bilinear branch at low
// resolution: the 1x1 convs commute with the interpolation).  out fp16 (or fp32) raster (B, rH, rW, Co).
int upsample_combine(const __half* Yp, const __half* Z, void* out, int out_f32, int B, int H, int W, int Co, int r,
                     cudaStream_t s);

// Folded tail: out[b][oc][y][x] = sum_{t=(dy,dx)} inb(y+dy-1, x+dx-1) * ( Qp[pix][oc*9+t] + bilinear4(Rb[..][oc*9+t])(pix) )
// Qp: fp32 [B*H*W*16][NT] rows ordered (b,h,w,i,j); Rb: fp32 [B*H*W][NT]; out fp32 NCHW (B, OC, 4H, 4W), or - out_fmt IMG_U8_NHWC -
// the 8-bit interleaved image rint(clamp(out, 0, 1) * 255) of demo.py:76-79 (torch.clamp + img_as_ubyte), (B, 4H, 4W, OC).
// With `ev` (validation forward, train.py:432-443; fp32 output only) the same pass also writes sigmoid(out) and accumulates
// the squared-error / Charbonnier sums against the target (3-channel targets are reduced to luminance when OC == 1, :437-438).
struct EvalEpilogue {
  const float* target = nullptr;  // (B, target_chans, 4H, 4W) fp32
  int target_chans = 0;           // OC, or 3 with OC == 1
  const float* weight = nullptr;  // (B, 1, 4H, 4W) fp32 or null (= ones)
  float* prob = nullptr;          // (B, OC, 4H, 4W) fp32 or null
  double* sums = nullptr;         // [5]: sum se, sum se*w, sum w, sum sqrt(d^2+eps^2)*w, element count; accumulated (caller zeroes)
  float eps = 1e-3f;
};
int tail_stencil(const float* Qp, const float* Rb, void* out, int out_fmt, const EvalEpilogue* ev, int B, int H, int W, int OC, int NT,
                 cudaStream_t s);

// ---- pre-pack helpers (run once per weight load)
// dst[n][k] = half(src[n][k] * (n < scale_rows ? scale : 1))
int pack_weight_f16(const float* src, __half* dst, int N, int K, int scale_rows, float scale, cudaStream_t s);
// dst [N][K + N] = [ half(src[N][K]) | identity ]  (residual folded into the GEMM as a second K segment)
int pack_weight_residual_f16(const float* src, __half* dst, int N, int K, cudaStream_t s);
// dst[perm(n)][k] = half(src[n][k]) with perm(c*rr + ij) = ij*Cq + c   (pixel-shuffle output reordering)
int pack_weight_shuffle_f16(const float* src, __half* dst, int Cq, int rr, int K, cudaStream_t s);
int scale_copy_f32(const float* src, float* dst, int n, int scale_n, float scale, cudaStream_t s);
// C[m][n] = sum_k A[m*lda + k] * B[k*ldb + n]   (tiny fp32 products for the linear folds)
int matmul_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int m, int n, int k, cudaStream_t s);
// 6x6 fold of conv_first and patch_embed.proj; wfold [Cin*36][E], bfold [E]
int fold_patch_embed(const float* w1, const float* b1, const float* w2, const float* b2, int Cin, int E, float* wfold,
                     float* bfold, cudaStream_t s);
// G[oc*9+t][c] = sum_m Wo[oc][m][t] * A[m][c] for t in 0..8; rows >= OC*9 zero.  Wo fp32 [OC][E][3][3], A fp32 [E][E]; G fp16 [NT][E]
int fold_tail_taps(const float* Wo, const float* A, __half* G, int OC, int E, int NT, cudaStream_t s);

}  // namespace sunet
