// tcgen05 GEMM for every linear layer / 1x1 conv on the SUNet forward path.
//   C[M,N] = epilogue( [A0 | A1][M, K0+K1] * W[N, K0+K1]^T )
// A0/A1/W are fp16, K-contiguous ("TN"); accumulation is fp32 in TMEM.
// Reference call sites this replaces: nn.Linear at SUNet_detail.py:13,15,99,101,298,652 and the
// 1x1 nn.Conv2d at SUNet_detail.py:343-363 (after NHWC re-interpretation).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

enum GemmAct { ACT_NONE = 0, ACT_GELU = 1, ACT_PRELU = 2 };

struct GemmArgs {
  const __half* A0 = nullptr;  // [M, K0], row stride lda0 (elements)
  int64_t lda0 = 0;
  int K0 = 0;
  const __half* A1 = nullptr;  // optional second K-segment (channel concat), [M, K1]
  int64_t lda1 = 0;
  int K1 = 0;
  const __half* W = nullptr;   // [N, K0+K1], row stride ldw
  int64_t ldw = 0;
  int64_t M = 0;
  int N = 0;
  const float* bias = nullptr;   // [N] fp32 or null
  int act = ACT_NONE;
  const float* prelu = nullptr;  // device scalar slope (ACT_PRELU)
  const __half* R = nullptr;     // residual [M, N] fp16 or null
  int64_t ldr = 0;
  void* C = nullptr;             // [M, N] fp16 (or fp32 when out_f32)
  int64_t ldc = 0;
  int out_f32 = 0;
  int force_block_n = 0;         // tuning / tests
  int pair_mode = 0;             // 0 auto, 1 single-CTA tiles, 2 CTA-pair (cta_group::2) 256-row tiles
  int dbg = 0;                   // bring-up ablations (timing only, wrong results): 1 no epilogue, 2 no MMA, 4 no TMA loads
};

struct GemmEpi {
  int64_t M;
  int N, K0, K1;
  int block_n, stages;
  const float* bias;
  const float* prelu;
  int act;
  const __half* R;
  int64_t ldr;
  void* C;
  int64_t ldc;
  int out_f32;
  int n_tiles;
  int64_t total_tiles;
  uint32_t acc_cols, tmem_cols;
  int dbg;
};

struct GemmOp {
  alignas(64) CUtensorMap tmA0;
  alignas(64) CUtensorMap tmA1;
  alignas(64) CUtensorMap tmW;
  GemmEpi epi;
  unsigned grid = 0;
  int smem = 0;
  int pair = 0;
  double flops = 0;
};

// returns 0 on success; on failure sets the thread-local error string (see error.h)
int gemm_prepare(const GemmArgs& a, GemmOp* op);
int gemm_launch(const GemmOp& op, cudaStream_t stream);
inline int gemm_run(const GemmArgs& a, cudaStream_t stream) {
  GemmOp op;
  int rc = gemm_prepare(a, &op);
  if (rc) return rc;
  return gemm_launch(op, stream);
}

// bring-up: one 128 x N x 64 UMMA from hand-swizzled shared memory (no TMA); D fp32 [128, N]
int umma_selftest(const __half* A, const __half* B, float* D, int N, cudaStream_t stream);

// encode a 2-D fp16 tensor map (inner dim contiguous), 128B swizzle, box {64, box_rows}
int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                     uint32_t box_rows);

}  // namespace sunet
