// Host-side engine + C ABI: parameter pre-pack (fp16 cast, qk_scale fold, linear folds) and the forward sequencing
// of SUNet (model/SUNet_detail.py:706-755) over the kernels in gemm_tcgen05.cu / attn_core.cu / elementwise.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sunet_b200.h"
#include "attn_core.cuh"
#include "attn_core_tc.cuh"
#include "attn_fused.cuh"
#include "elementwise.cuh"
#include "error.h"
#include "gemm.cuh"
#include "mlp_fused.cuh"
#include "mlp_row.cuh"
#include "proj_ln.cuh"
#include "tail_fused.cuh"

namespace sunet {

// ------------------------------------------------------------------------------------------------ infrastructure
struct Arena {  // device allocations owned by a handle (pre-pack time only)
  std::vector<void*> ptrs;
  ~Arena() { for (void* p : ptrs) cudaFree(p); }
  int alloc(void** out, size_t bytes) {
    void* p = nullptr;
    SUNET_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    ptrs.push_back(p);
    *out = p;
    return 0;
  }
  template <typename T> int alloc_t(T** out, size_t n) { return alloc(reinterpret_cast<void**>(out), n * sizeof(T)); }
};

struct Params {
  std::unordered_map<std::string, std::pair<const float*, int64_t>> m;
  bool has(const std::string& k) const { return m.count(k) != 0; }
  int get(const std::string& k, int64_t numel, const float** out) const {
    auto it = m.find(k);
    if (it == m.end()) return fail(SUNET_E_ARG, "missing parameter '%s'", k.c_str());
    if (it->second.second != numel)
      return fail(SUNET_E_SHAPE, "parameter '%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.second, (long long)numel);
    *out = it->second.first;
    return 0;
  }
};

// bump allocator over caller-provided scratch; `dry` only measures the high-water mark
struct Scratch {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  int take(void** out, size_t bytes) {
    const size_t a = (off + 255) & ~size_t(255);
    if (!dry && a + bytes > cap) return fail(SUNET_E_WORKSPACE, "workspace too small: need %zu bytes, have %zu", a + bytes, cap);
    *out = base + a;
    off = a + bytes;
    if (off > peak) peak = off;
    return 0;
  }
  template <typename T> int take_t(T** out, size_t n) { return take(reinterpret_cast<void**>(out), n * sizeof(T)); }
};
struct ScratchMark {  // stack discipline
  Scratch& s; size_t saved;
  explicit ScratchMark(Scratch& sc) : s(sc), saved(sc.off) {}
  ~ScratchMark() { s.off = saved; }
};

enum KernelKind { K_GEMM = 0, K_ATTN = 1, K_LAYERNORM = 2, K_MERGE_LN = 3, K_PATCH_EMBED = 4, K_UP_COMBINE = 5, K_TAIL = 6, K_CAST = 7, K_IM2COL = 8, K_MLP_FUSED = 9, K_ATTN_FUSED = 10, K_TAIL_FUSED = 11 };

struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t e0, e1;
};

struct Ctx {
  Scratch sc;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  std::vector<ProfRec>* prof = nullptr;  // when set, every launch is bracketed by CUDA events on `stream`
  // image formats at the two ends of the whole-model forward, and the optional validation epilogue of the last kernel
  int in_fmt = IMG_F32_NCHW, out_fmt = IMG_F32_NCHW;
  const EvalEpilogue* ev = nullptr;
  bool dry() const { return sc.dry; }
};

static int prof_begin(Ctx& c, int kind, double flops, double bytes) {
  ProfRec r{kind, flops, bytes, nullptr, nullptr};
  SUNET_CUDA(cudaEventCreate(&r.e0));
  SUNET_CUDA(cudaEventCreate(&r.e1));
  SUNET_CUDA(cudaEventRecord(r.e0, c.stream));
  c.prof->push_back(r);
  return 0;
}
static int prof_end(Ctx& c) {
  SUNET_CUDA(cudaEventRecord(c.prof->back().e1, c.stream));
  return 0;
}

// one kernel launch: counted always, executed unless this is a dry (sizing) pass, event-bracketed when profiling.
// flops / bytes are the ALGORITHMIC figures of the launch (2*M*N*K; each operand read once, result written once).
#define RUN(ctx, kind, flops, bytes, expr)                         \
  do {                                                             \
    (ctx).launches++;                                              \
    if (!(ctx).dry()) {                                            \
      if ((ctx).prof) SUNET_TRY(prof_begin((ctx), (kind), (flops), (bytes))); \
      SUNET_TRY(expr);                                             \
      if ((ctx).prof) SUNET_TRY(prof_end((ctx)));                  \
    }                                                              \
  } while (0)

struct Linear {
  __half* w = nullptr;
  float* b = nullptr;
  int N = 0, K = 0;
  int res_fold = 0;   // 1: the last N columns of w are an identity block (residual folded in as a K segment)
};

static int pack_linear(Arena& ar, const Params& P, const std::string& wkey, const std::string& bkey, int N, int K, Linear* L,
                       cudaStream_t s, int scale_rows = 0, float scale = 1.f) {
  const float* w;
  SUNET_TRY(P.get(wkey, static_cast<int64_t>(N) * K, &w));
  L->N = N; L->K = K;
  SUNET_TRY(ar.alloc_t(&L->w, static_cast<size_t>(N) * K));
  SUNET_TRY(pack_weight_f16(w, L->w, N, K, scale_rows, scale, s));
  if (!bkey.empty() && P.has(bkey)) {
    const float* b;
    SUNET_TRY(P.get(bkey, N, &b));
    SUNET_TRY(ar.alloc_t(&L->b, N));
    SUNET_TRY(scale_copy_f32(b, L->b, N, scale_rows, scale, s));
  }
  return 0;
}
static int copy_vec(Arena& ar, const Params& P, const std::string& key, int n, float** out, cudaStream_t s) {
  const float* src;
  SUNET_TRY(P.get(key, n, &src));
  SUNET_TRY(ar.alloc_t(out, n));
  SUNET_CUDA(cudaMemcpyAsync(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

static int run_linear(Ctx& c, const Linear& L, const __half* A, int64_t lda, int64_t M, void* C, int64_t ldc, int act = ACT_NONE,
                      const float* prelu = nullptr, const __half* R = nullptr, int64_t ldr = 0, int out_f32 = 0,
                      const __half* A1 = nullptr, int64_t lda1 = 0, int K1 = 0) {
  GemmArgs a;
  a.A0 = A; a.lda0 = lda; a.K0 = L.K - K1;
  a.A1 = A1; a.lda1 = lda1; a.K1 = K1;
  a.W = L.w; a.ldw = L.K; a.M = M; a.N = L.N;
  a.bias = L.b; a.act = act; a.prelu = prelu; a.R = R; a.ldr = ldr; a.C = C; a.ldc = ldc; a.out_f32 = out_f32;
  if (L.res_fold) a.force_block_n = L.N;   // one N tile per row block: a row's shortcut columns are only read by the tile that writes them
  // algorithmic figures: the identity block of a folded residual is not counted as FLOPs, its operand is the residual read
  const double K = L.res_fold ? L.K - L.N : L.K;
  const double bytes = 2.0 * M * K + 2.0 * L.N * K + (out_f32 ? 4.0 : 2.0) * M * L.N + ((R || L.res_fold) ? 2.0 * M * L.N : 0.0);
  RUN(c, K_GEMM, 2.0 * M * L.N * K, bytes, gemm_run(a, c.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ module packs
struct Handle {
  std::string kind;
  Arena arena;
  virtual ~Handle() {}
};

struct AttnPack {  // WindowAttention parameters (SUNet_detail.py:83-105)
  int dim = 0, heads = 0;
  float scale = 1.f;
  Linear qkv, proj;
  Linear proj_res;   // [proj.weight | I]: proj + shortcut as one two-segment GEMM (dims whose output is a single N tile)
  float* table = nullptr;
  int pack(Arena& ar, const Params& P, const std::string& pre, int dim_, int heads_, double qk_scale, cudaStream_t s) {
    dim = dim_; heads = heads_;
    if (heads <= 0 || dim % heads) return fail(SUNET_E_SHAPE, "attention: dim %d not divisible by heads %d", dim, heads);
    scale = qk_scale > 0 ? static_cast<float>(qk_scale) : 1.f / sqrtf(static_cast<float>(dim / heads));  // :80
    // q = q * scale (:117) folded into the first `dim` rows of qkv.weight / qkv.bias, together with log2(e): the
    // attention core works in the exp2 domain (its bias table / mask are scaled the same way when staged)
    SUNET_TRY(pack_linear(ar, P, pre + "qkv.weight", pre + "qkv.bias", 3 * dim, dim, &qkv, s, dim, scale * 1.4426950408889634f));
    SUNET_TRY(pack_linear(ar, P, pre + "proj.weight", pre + "proj.bias", dim, dim, &proj, s));
    if (dim <= 256 && dim % 16 == 0 && getenv("SUNET_NO_RESIDUAL_FOLD") == nullptr) {
      const float* w;
      SUNET_TRY(P.get(pre + "proj.weight", static_cast<int64_t>(dim) * dim, &w));
      proj_res.N = dim; proj_res.K = 2 * dim; proj_res.b = proj.b; proj_res.res_fold = 1;
      SUNET_TRY(ar.alloc_t(&proj_res.w, static_cast<size_t>(dim) * 2 * dim));
      SUNET_TRY(pack_weight_residual_f16(w, proj_res.w, dim, dim, s));
    }
    SUNET_TRY(copy_vec(ar, P, pre + "relative_position_bias_table", 225 * heads, &table, s));
    return 0;
  }
};

struct MlpPack {
  int cin = 0, hid = 0, cout = 0;
  Linear fc1, fc2;
  int pack(Arena& ar, const Params& P, const std::string& pre, int cin_, int hid_, int cout_, cudaStream_t s) {
    cin = cin_; hid = hid_; cout = cout_;
    SUNET_TRY(pack_linear(ar, P, pre + "fc1.weight", pre + "fc1.bias", hid, cin, &fc1, s));
    SUNET_TRY(pack_linear(ar, P, pre + "fc2.weight", pre + "fc2.bias", cout, hid, &fc2, s));
    return 0;
  }
};

struct BlockPack {  // SwinTransformerBlock (SUNet_detail.py:176-225)
  int dim = 0, H = 0, W = 0, heads = 0, shift = 0;
  float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  AttnPack attn;
  MlpPack mlp;
  MlpFusedPack mf;      // norm2 + mlp + residual as one kernel (dims with a fused instantiation)
  bool use_mf = false;
  AttnFusedPack af;     // norm1 + shift/partition + qkv + attention core + reverse/un-shift as one kernel
  bool use_af = false;
  float* bias_exp = nullptr;   // [heads][64][64] relative-position bias, times log2(e): the tcgen05 attention core reads whole rows
  bool use_tc_core = false;    // head_dim 96 on a one-window grid (stage 3): QK^T and PV on tcgen05 (attn_core_tc.cu)
  bool use_proj_ln = false, use_row_gemm = false;   // C = 384 whole-row kernels (proj + shortcut + norm2; fc2 + residual)
  MlpRowPack mr;        // C = 384: fc1 + GELU + fc2 + residual as one whole-row kernel (the hidden activation stays on the SM)
  bool use_row_mlp = false;
  int64_t row_min_m = 2048;   // rows below which the C = 384 back half runs as proj_ln + GEMMs (SUNET_ROW_MLP_MIN_M at pre-pack: tests force 0)
  int pack(Arena& ar, const Params& P, const std::string& pre, int dim_, int H_, int W_, int heads_, int shift_, double qk_scale,
           cudaStream_t s) {
    dim = dim_; H = H_; W = W_; heads = heads_; shift = shift_;
    if (H % 8 || W % 8) return fail(SUNET_E_SHAPE, "swin block: grid %dx%d must be a multiple of the 8x8 window", H, W);
    if ((H < W ? H : W) <= 8) shift = 0;  // :186-189
    if (shift != 0 && shift != 4) return fail(SUNET_E_SHAPE, "swin block: shift_size %d unsupported (0 or 4)", shift);
    SUNET_TRY(copy_vec(ar, P, pre + "norm1.weight", dim, &g1, s));
    SUNET_TRY(copy_vec(ar, P, pre + "norm1.bias", dim, &b1, s));
    SUNET_TRY(copy_vec(ar, P, pre + "norm2.weight", dim, &g2, s));
    SUNET_TRY(copy_vec(ar, P, pre + "norm2.bias", dim, &b2, s));
    SUNET_TRY(attn.pack(ar, P, pre + "attn.", dim, heads, qk_scale, s));
    use_af = attn_fused_supported(dim, heads) && getenv("SUNET_NO_FUSED_ATTN") == nullptr;
    if (use_af) {
      const float *gw, *gb, *wq, *bq = nullptr;
      SUNET_TRY(P.get(pre + "norm1.weight", dim, &gw));
      SUNET_TRY(P.get(pre + "norm1.bias", dim, &gb));
      SUNET_TRY(P.get(pre + "attn.qkv.weight", static_cast<int64_t>(3) * dim * dim, &wq));
      if (P.has(pre + "attn.qkv.bias")) SUNET_TRY(P.get(pre + "attn.qkv.bias", 3 * dim, &bq));
      SUNET_TRY(ar.alloc_t(&af.w, static_cast<size_t>(3) * dim * attn_fused_w_pitch(dim)));
      SUNET_TRY(ar.alloc_t(&af.hconst, static_cast<size_t>(6) * dim));
      SUNET_TRY(attn_fused_prepack(&af, dim, heads, attn.scale * 1.4426950408889634f, gw, gb, wq, bq, attn.table, s));
    }
    use_tc_core = !use_af && attn_core_tc_supported(dim, heads, H, W, shift) && getenv("SUNET_NO_TC_CORE") == nullptr;
    if (use_tc_core) {
      SUNET_TRY(ar.alloc_t(&bias_exp, static_cast<size_t>(heads) * 4096));
      SUNET_TRY(attn_core_tc_expand_bias(attn.table, heads, bias_exp, s));
    }
    use_mf = mlp_fused_supported(dim) && getenv("SUNET_NO_FUSED_MLP") == nullptr;
    if (use_mf) {
      const float *gw, *gb, *w1, *b1, *w2, *b2v;
      SUNET_TRY(P.get(pre + "norm2.weight", dim, &gw));
      SUNET_TRY(P.get(pre + "norm2.bias", dim, &gb));
      SUNET_TRY(P.get(pre + "mlp.fc1.weight", static_cast<int64_t>(4) * dim * dim, &w1));
      SUNET_TRY(P.get(pre + "mlp.fc1.bias", 4 * dim, &b1));
      SUNET_TRY(P.get(pre + "mlp.fc2.weight", static_cast<int64_t>(4) * dim * dim, &w2));
      SUNET_TRY(P.get(pre + "mlp.fc2.bias", dim, &b2v));
      SUNET_TRY(ar.alloc_t(&mf.w1g, static_cast<size_t>(4) * dim * dim));
      SUNET_TRY(ar.alloc_t(&mf.w2, static_cast<size_t>(4) * dim * dim));
      SUNET_TRY(ar.alloc_t(&mf.hconst, static_cast<size_t>(8) * dim));
      SUNET_TRY(ar.alloc_t(&mf.b2, dim));
      SUNET_TRY(mlp_fused_prepack(&mf, dim, gw, gb, w1, b1, w2, b2v, s));
      if (getenv("SUNET_NO_FUSED_PROJ") == nullptr) {   // attn.proj + first residual ride in front of the MLP kernel
        const float *wp, *bp = nullptr;
        SUNET_TRY(P.get(pre + "attn.proj.weight", static_cast<int64_t>(dim) * dim, &wp));
        if (P.has(pre + "attn.proj.bias")) SUNET_TRY(P.get(pre + "attn.proj.bias", dim, &bp));
        SUNET_TRY(ar.alloc_t(&mf.wp, static_cast<size_t>(dim) * dim));
        SUNET_TRY(ar.alloc_t(&mf.bp, dim));
        SUNET_TRY(ar.alloc_t(&mf.w1h, static_cast<size_t>(4) * dim * mlp_fused_w1h_pitch(dim)));
        SUNET_TRY(ar.alloc_t(&mf.hbias, static_cast<size_t>(4) * dim));
        SUNET_TRY(mlp_fused_set_proj(&mf, wp, bp, gw, gb, w1, b1, s));
      }
    } else {
      SUNET_TRY(mlp.pack(ar, P, pre + "mlp.", dim, 4 * dim, dim, s));
    }
    // kernel selection is fixed here, once per pre-pack: the forward never consults the environment
    use_proj_ln = !use_mf && proj_ln_supported(dim) && getenv("SUNET_NO_PROJ_LN") == nullptr;
    use_row_gemm = row_gemm_supported(dim, 4 * dim) && getenv("SUNET_NO_ROW_GEMM") == nullptr;
    use_row_mlp = use_proj_ln && mlp_row_supported(dim) && getenv("SUNET_NO_ROW_MLP") == nullptr;
    if (const char* e = getenv("SUNET_ROW_MLP_MIN_M")) row_min_m = atoll(e);
    if (use_row_mlp) {
      Linear fc1h;   // 0.5 * fc1 (weights and bias): the GELU epilogue of the whole-row kernel takes u = x / 2
      SUNET_TRY(pack_linear(ar, P, pre + "mlp.fc1.weight", pre + "mlp.fc1.bias", 4 * dim, dim, &fc1h, s, 4 * dim, 0.5f));
      if (!fc1h.b) {
        SUNET_TRY(ar.alloc_t(&fc1h.b, static_cast<size_t>(4) * dim));
        SUNET_CUDA(cudaMemsetAsync(fc1h.b, 0, static_cast<size_t>(4) * dim * sizeof(float), s));
      }
      mr.C = dim; mr.w1h = fc1h.w; mr.hbias = fc1h.b; mr.w2 = mlp.fc2.w; mr.b2 = mlp.fc2.b;
      SUNET_TRY(mlp_row_prepare(&mr));
      if (getenv("SUNET_NO_ROW_PROJ") == nullptr) {   // attn.proj + first residual + norm2 ride in front of the same kernel
        const float *gw, *gb, *w1, *b1v = nullptr;
        SUNET_TRY(P.get(pre + "norm2.weight", dim, &gw));
        SUNET_TRY(P.get(pre + "norm2.bias", dim, &gb));
        SUNET_TRY(P.get(pre + "mlp.fc1.weight", static_cast<int64_t>(4) * dim * dim, &w1));
        if (P.has(pre + "mlp.fc1.bias")) SUNET_TRY(P.get(pre + "mlp.fc1.bias", 4 * dim, &b1v));
        __half* w1hg; float* hbg;
        SUNET_TRY(ar.alloc_t(&w1hg, static_cast<size_t>(4) * dim * dim));
        SUNET_TRY(ar.alloc_t(&hbg, static_cast<size_t>(8) * dim));
        SUNET_TRY(mlp_row_set_proj(&mr, attn.proj.w, attn.proj.b, gw, gb, w1, b1v, w1hg, hbg, s));
      }
    }
    return 0;
  }
  // The window-attention part of the block up to the per-head attention output O (norm1, shift, partition, qkv, QK^T + bias +
  // mask, softmax, AV, reverse, un-shift; SUNet_detail.py:233-257 without the proj Linear of :136): x_in -> O [B*H*W][dim]
  int attention_part(Ctx& c, const __half* x_in, __half* O, __half* T, __half* QKV, int B) const {
    const int64_t M = static_cast<int64_t>(B) * H * W;
    if (use_af) {
      RUN(c, K_ATTN_FUSED, 6.0 * M * dim * dim + 256.0 * M * dim, 4.0 * M * dim,
          attn_fused_launch(af, x_in, O, B, H, W, shift, c.stream));                             // :233-257 minus proj
    } else {
      RUN(c, K_LAYERNORM, 0.0, 4.0 * M * dim, layernorm_f16(x_in, dim, T, dim, g1, b1, M, dim, c.stream));                     // :233
      SUNET_TRY(run_linear(c, attn.qkv, T, dim, M, QKV, 3 * dim));                             // :114
      if (use_tc_core) {   // one window per image, no shift: q/k/v tiles by TMA, S and O in TMEM
        RUN(c, K_ATTN, 256.0 * M * dim, 8.0 * M * dim, attn_core_tc_launch(QKV, 3 * dim, O, dim, M, dim, heads, bias_exp, c.stream));   // :118-135
        return 0;
      }
      AttnCoreArgs a;
      a.qkv = QKV; a.ld = 3 * dim; a.out = O; a.ldo = dim; a.B = B; a.H = H; a.W = W; a.C = dim; a.heads = heads;
      a.shift = shift; a.bias_table = attn.table; a.mask_mode = shift > 0 ? 1 : 0;
      RUN(c, K_ATTN, 256.0 * M * dim, 8.0 * M * dim, attn_core_launch(a, c.stream));             // :118-135, :236-257
    }
    return 0;
  }
  // x_in [B*H*W][dim] -> x_out (may alias x_in); image-order rows throughout
  int forward(Ctx& c, const __half* x_in, __half* x_out, int B) const {
    ScratchMark mk(c.sc);
    const int64_t M = static_cast<int64_t>(B) * H * W;
    __half *T, *QKV, *O, *Hd;
    // The whole-row kernel runs one tile per CTA pair, twelve hidden chunks in sequence: ~40 us however few rows there are.  Below
    // 2048 rows (batch < 8 at stage 2) the proj_ln + two-GEMM path, which spreads a small M over all SMs, has the lower latency
    // (batch 1: 2.27 -> 2.16 ms per forward, profiles/r06_latency.json).
    const bool row_mlp = use_row_mlp && M >= row_min_m;
    SUNET_TRY(c.sc.take_t(&T, (use_af && (use_mf || (row_mlp && mr.has_proj))) ? 0 : M * dim));
    SUNET_TRY(c.sc.take_t(&QKV, use_af ? 0 : M * 3 * dim));
    SUNET_TRY(c.sc.take_t(&O, M * dim));
    SUNET_TRY(c.sc.take_t(&Hd, (use_mf || row_mlp) ? 0 : M * 4 * dim));
    SUNET_TRY(attention_part(c, x_in, O, T, QKV, B));
    if (use_mf && mf.has_proj) {   // :136 proj, :261 shortcut add, :262 norm2 + Mlp + residual: one kernel
      RUN(c, K_MLP_FUSED, 18.0 * M * dim * dim, 6.0 * M * dim, mlp_proj_fused_launch(mf, O, x_in, x_out, M, c.stream));
      return 0;
    }
    if (row_mlp && mr.has_proj) {   // :136 proj, :261 shortcut add, :262 norm2 + Mlp + residual: one CTA-pair whole-row kernel
      RUN(c, K_MLP_FUSED, 18.0 * M * dim * dim, 6.0 * M * dim + 18.0 * dim * dim, mlp_row_proj_launch(mr, O, x_in, x_out, M, c.stream));
      return 0;
    }
    if (use_proj_ln) {
      // :136 proj, :261 shortcut, :262 norm2 in one kernel (whole rows per CTA); then the two MLP GEMMs
      ProjLnPack pl;
      pl.w = attn.proj.w; pl.bias = attn.proj.b; pl.gamma = g2; pl.beta = b2; pl.C = dim;
      RUN(c, K_GEMM, 2.0 * M * dim * dim, 8.0 * M * dim + 2.0 * dim * dim, proj_ln_launch(pl, O, x_in, x_out, T, M, c.stream));
      if (row_mlp) {   // :19-22, :262: fc1 + GELU + fc2 + second residual, the hidden activation never leaves the SM
        RUN(c, K_MLP_FUSED, 16.0 * M * dim * dim, 6.0 * M * dim + 16.0 * dim * dim, mlp_row_launch(mr, T, x_out, x_out, M, c.stream));
        return 0;
      }
      SUNET_TRY(run_linear(c, mlp.fc1, T, dim, M, Hd, 4 * dim, ACT_GELU));                       // :19-20
      if (use_row_gemm) {
        RUN(c, K_GEMM, 8.0 * M * dim * dim, 12.0 * M * dim + 8.0 * dim * dim,
            row_gemm_residual_launch(mlp.fc2.w, mlp.fc2.b, dim, 4 * dim, Hd, x_out, x_out, M, c.stream));   // :22, :262
      } else {
        SUNET_TRY(run_linear(c, mlp.fc2, Hd, 4 * dim, M, x_out, dim, ACT_NONE, nullptr, x_out, dim));  // :22, :262
      }
      return 0;
    }
    if (attn.proj_res.w) {   // :136, :261 - the shortcut rides the TMA ring as a second K segment against an identity block
      SUNET_TRY(run_linear(c, attn.proj_res, O, dim, M, x_out, dim, ACT_NONE, nullptr, nullptr, 0, 0, x_in, dim, dim));
    } else {
      SUNET_TRY(run_linear(c, attn.proj, O, dim, M, x_out, dim, ACT_NONE, nullptr, x_in, dim));   // :136, :261
    }
    if (use_mf) {
      RUN(c, K_MLP_FUSED, 16.0 * M * dim * dim, 4.0 * M * dim, mlp_fused_launch(mf, x_out, x_out, M, c.stream));               // :262 (norm2, mlp, +res)
      return 0;
    }
    RUN(c, K_LAYERNORM, 0.0, 4.0 * M * dim, layernorm_f16(x_out, dim, T, dim, g2, b2, M, dim, c.stream));                      // :262
    SUNET_TRY(run_linear(c, mlp.fc1, T, dim, M, Hd, 4 * dim, ACT_GELU));                       // :19-20
    SUNET_TRY(run_linear(c, mlp.fc2, Hd, 4 * dim, M, x_out, dim, ACT_NONE, nullptr, x_out, dim));  // :22, :262
    return 0;
  }
};

struct MergePack {  // PatchMerging (SUNet_detail.py:294-322)
  int dim = 0, H = 0, W = 0;
  float *g = nullptr, *b = nullptr;
  Linear red;
  int pack(Arena& ar, const Params& P, const std::string& pre, int dim_, int H_, int W_, cudaStream_t s) {
    dim = dim_; H = H_; W = W_;
    SUNET_TRY(copy_vec(ar, P, pre + "norm.weight", 4 * dim, &g, s));
    SUNET_TRY(copy_vec(ar, P, pre + "norm.bias", 4 * dim, &b, s));
    SUNET_TRY(pack_linear(ar, P, pre + "reduction.weight", "", 2 * dim, 4 * dim, &red, s));
    return 0;
  }
  int forward(Ctx& c, const __half* x, __half* y, int B) const {
    ScratchMark mk(c.sc);
    const int64_t M4 = static_cast<int64_t>(B) * (H / 2) * (W / 2);
    __half* T;
    SUNET_TRY(c.sc.take_t(&T, M4 * 4 * dim));
    RUN(c, K_MERGE_LN, 0.0, 16.0 * M4 * dim, merge_gather_ln_f16(x, T, g, b, B, H, W, dim, c.stream));
    SUNET_TRY(run_linear(c, red, T, 4 * dim, M4, y, 2 * dim));
    return 0;
  }
};

// Dual up-sample (SUNet_detail.py:335-386) with its linear tail folded:
//   out = conv([up_p, up_b]) = (Wc[:, :Cq] Wp3) PS(prelu(Wp0 x)) + bilinear((Wc[:, Cq:] Wb3) prelu(Wb0 x + b))
// (1x1 convs without bias commute with the interpolation, whose taps sum to 1).
struct UpPack {
  int C = 0, r = 0, H = 0, W = 0, Cq = 0, Co = 0;
  Linear p0, b0, pp, zz;
  float *slope_p = nullptr, *slope_b = nullptr;
  float *Ap = nullptr, *Ab = nullptr;  // fp32 folded matrices (kept for the tail fold)
  int pack(Arena& ar, const Params& P, const std::string& pre, int C_, int r_, int H_, int W_, cudaStream_t s) {
    C = C_; r = r_; H = H_; W = W_;
    if (r != 2 && r != 4) return fail(SUNET_E_SHAPE, "upsample: factor %d unsupported (2 or 4)", r);
    Cq = r == 2 ? C / 2 : C;
    Co = Cq;
    const int rr = r * r;
    const float *wc, *wp0, *wp3, *wb3;
    SUNET_TRY(P.get(pre + "conv.weight", static_cast<int64_t>(Co) * 2 * Cq, &wc));
    SUNET_TRY(P.get(pre + "up_p.0.weight", static_cast<int64_t>(rr) * Cq * C, &wp0));
    SUNET_TRY(P.get(pre + "up_p.3.weight", static_cast<int64_t>(Cq) * Cq, &wp3));
    SUNET_TRY(P.get(pre + "up_b.3.weight", static_cast<int64_t>(Co) * C, &wb3));
    // pixel-shuffle reorder of up_p[0] rows: (c*rr + ij) -> (ij*Cq + c), so each hi-res pixel's Cq channels are contiguous
    p0.N = rr * Cq; p0.K = C;
    SUNET_TRY(ar.alloc_t(&p0.w, static_cast<size_t>(p0.N) * C));
    SUNET_TRY(pack_weight_shuffle_f16(wp0, p0.w, Cq, rr, C, s));
    SUNET_TRY(pack_linear(ar, P, pre + "up_b.0.weight", pre + "up_b.0.bias", C, C, &b0, s));
    SUNET_TRY(copy_vec(ar, P, pre + "up_p.1.weight", 1, &slope_p, s));
    SUNET_TRY(copy_vec(ar, P, pre + "up_b.1.weight", 1, &slope_b, s));
    SUNET_TRY(ar.alloc_t(&Ap, static_cast<size_t>(Co) * Cq));
    SUNET_TRY(ar.alloc_t(&Ab, static_cast<size_t>(Co) * C));
    SUNET_TRY(matmul_f32(wc, 2 * Cq, wp3, Cq, Ap, Cq, Co, Cq, Cq, s));            // Wc[:, :Cq] @ Wp3
    SUNET_TRY(matmul_f32(wc + Cq, 2 * Cq, wb3, C, Ab, C, Co, C, Co, s));          // Wc[:, Cq:] @ Wb3
    pp.N = Co; pp.K = Cq; zz.N = Co; zz.K = C;
    SUNET_TRY(ar.alloc_t(&pp.w, static_cast<size_t>(Co) * Cq));
    SUNET_TRY(ar.alloc_t(&zz.w, static_cast<size_t>(Co) * C));
    SUNET_TRY(pack_weight_f16(Ap, pp.w, Co, Cq, 0, 1.f, s));
    SUNET_TRY(pack_weight_f16(Ab, zz.w, Co, C, 0, 1.f, s));
    return 0;
  }
  // x [B*H*W][C] -> out raster [B*rH*rW][Co] (fp16, or fp32 when out_f32)
  int forward(Ctx& c, const __half* x, void* out, int out_f32, int B) const {
    ScratchMark mk(c.sc);
    const int64_t M = static_cast<int64_t>(B) * H * W;
    const int rr = r * r;
    __half *Pb, *Yp, *Bb, *Z;
    SUNET_TRY(c.sc.take_t(&Pb, M * rr * Cq));
    SUNET_TRY(c.sc.take_t(&Yp, M * rr * Co));
    SUNET_TRY(c.sc.take_t(&Bb, M * C));
    SUNET_TRY(c.sc.take_t(&Z, M * Co));
    SUNET_TRY(run_linear(c, p0, x, C, M, Pb, rr * Cq, ACT_PRELU, slope_p));           // up_p[0..2]
    SUNET_TRY(run_linear(c, pp, Pb, Cq, M * rr, Yp, Co));                             // up_p[3] + conv (p half)
    SUNET_TRY(run_linear(c, b0, x, C, M, Bb, C, ACT_PRELU, slope_b));                 // up_b[0..1]
    SUNET_TRY(run_linear(c, zz, Bb, C, M, Z, Co));                                    // up_b[3] + conv (b half), at low res
    RUN(c, K_UP_COMBINE, 0.0, (2.0 * rr + 2.0 + (out_f32 ? 4.0 : 2.0) * rr) * M * Co, upsample_combine(Yp, Z, out, out_f32, B, H, W, Co, r, c.stream));          // pixel shuffle + bilinear + add
    return 0;
  }
};

// Final x4 Dual up-sample + 3x3 output conv (SUNet_detail.py:742-753) folded into per-tap scalar maps.
struct TailPack {
  UpPack up;
  int E = 0, OC = 0, NT = 0, H = 0, W = 0;
  bool fused = false;   // tail_up_fused instead of two GEMMs (decided at pre-pack)
  Linear gp, gb;
  int pack(Arena& ar, const Params& P, const std::string& up_pre, const std::string& out_key, int E_, int OC_, int H_, int W_,
           cudaStream_t s) {
    E = E_; OC = OC_; H = H_; W = W_;
    if (OC < 1 || OC > 3) return fail(SUNET_E_SHAPE, "out_chans %d unsupported (1..3)", OC);
    NT = (OC * 9 + 15) / 16 * 16;
    SUNET_TRY(up.pack(ar, P, up_pre, E, 4, H, W, s));
    const float* wo;
    SUNET_TRY(P.get(out_key, static_cast<int64_t>(OC) * E * 9, &wo));
    gp.N = NT; gp.K = E; gb.N = NT; gb.K = E;
    SUNET_TRY(ar.alloc_t(&gp.w, static_cast<size_t>(NT) * E));
    SUNET_TRY(ar.alloc_t(&gb.w, static_cast<size_t>(NT) * E));
    SUNET_TRY(fold_tail_taps(wo, up.Ap, gp.w, OC, E, NT, s));
    SUNET_TRY(fold_tail_taps(wo, up.Ab, gb.w, OC, E, NT, s));
    fused = tail_up_fused_supported(E, NT, W) && getenv("SUNET_NO_FUSED_TAIL") == nullptr;
    return 0;
  }
  int forward(Ctx& c, const __half* x, void* out, int B) const {
    ScratchMark mk(c.sc);
    const int64_t M = static_cast<int64_t>(B) * H * W;
    __half *Pb, *Bb;
    float *Qp, *Rb;
    SUNET_TRY(c.sc.take_t(&Pb, fused ? 0 : M * 16 * E));
    SUNET_TRY(c.sc.take_t(&Qp, fused ? tail_strips_floats(M) : M * 16 * NT));
    SUNET_TRY(c.sc.take_t(&Bb, M * E));
    SUNET_TRY(c.sc.take_t(&Rb, M * NT));
    SUNET_TRY(run_linear(c, up.b0, x, E, M, Bb, E, ACT_PRELU, up.slope_b));
    SUNET_TRY(run_linear(c, gb, Bb, E, M, Rb, NT, ACT_NONE, nullptr, nullptr, 0, 1));
    if (fused) {
      // up_p[0..1], the folded taps, the bilinear branch and the per-token part of the 3x3 stencil in one kernel: neither the
      // [M][16 * 96] activation nor the [M * 16][16] tap values reach HBM, only 96 floats of partial output sums per token
      RUN(c, K_TAIL_FUSED, 2.0 * M * 16 * E * E + 2.0 * M * 16 * NT * E, 2.0 * M * E + 4.0 * M * NT + 4.0 * M * 96,
          tail_up_fused_launch(x, up.p0.w, gp.w, up.slope_p, Rb, Qp, B, H, W, c.stream));
      RUN(c, K_TAIL, 0.0, 4.0 * M * 96 + (c.out_fmt == IMG_U8_NHWC ? 1.0 : 4.0) * M * 16 * OC, tail_finish(Qp, out, c.out_fmt, c.ev, B, H, W, c.stream));
      return 0;
    }
    SUNET_TRY(run_linear(c, up.p0, x, E, M, Pb, 16 * E, ACT_PRELU, up.slope_p));
    SUNET_TRY(run_linear(c, gp, Pb, E, M * 16, Qp, NT, ACT_NONE, nullptr, nullptr, 0, 1));
    RUN(c, K_TAIL, 0.0, 4.0 * M * NT * 17 + (c.out_fmt == IMG_U8_NHWC ? 1.0 : 4.0) * M * 16 * OC,
        tail_stencil(Qp, Rb, out, c.out_fmt, c.ev, B, H, W, OC, NT, c.stream));
    return 0;
  }
};

struct PatchEmbedPack {  // stand-alone PatchEmbed (SUNet_detail.py:529-556)
  int cin = 0, E = 0, P = 0, has_norm = 0;
  Linear proj;
  float *g = nullptr, *b = nullptr;
};

struct ModelPack {
  int img = 0, patch = 0, in_chans = 0, out_chans = 0, E = 0, G = 0;
  int depths[4] = {0, 0, 0, 0}, heads[4] = {0, 0, 0, 0};
  float *wfold = nullptr, *bfold = nullptr, *pe_g = nullptr, *pe_b = nullptr;
  __half* wpk = nullptr;   // wfold as fp16 [E][PATCH_EMBED_WPK_PITCH] for the tensor-core patch-embed kernel
  std::vector<BlockPack> enc[4], dec[4];
  MergePack merge[3];
  float *norm_g = nullptr, *norm_b = nullptr, *normup_g = nullptr, *normup_b = nullptr;
  UpPack up0, ups[4];
  Linear cat[4];
  TailPack tail;
};

// ------------------------------------------------------------------------------------------------ typed handles
struct BlockHandle : Handle { BlockPack p; };
struct AttnHandle : Handle { AttnPack p; };
struct MlpHandle : Handle { MlpPack p; };
struct MergeHandle : Handle { MergePack p; };
struct UpHandle : Handle { UpPack p; };
struct PatchEmbedHandle : Handle { PatchEmbedPack p; };
struct ModelHandle : Handle { ModelPack p; };

static int pack_model(ModelHandle* h, const Params& P, const int64_t* ia, int nia, double qk_scale, cudaStream_t s) {
  if (nia < 14) return fail(SUNET_E_ARG, "sunet: 14 integer arguments expected, got %d", nia);
  ModelPack& m = h->p;
  Arena& ar = h->arena;
  m.img = (int)ia[0]; m.patch = (int)ia[1]; m.in_chans = (int)ia[2]; m.out_chans = (int)ia[3]; m.E = (int)ia[4];
  const int window = (int)ia[5];
  for (int i = 0; i < 4; ++i) { m.depths[i] = (int)ia[6 + i]; m.heads[i] = (int)ia[10 + i]; }
  if (window != 8) return fail(SUNET_E_SHAPE, "sunet: window_size %d unsupported (8)", window);
  if (m.patch != 4) return fail(SUNET_E_SHAPE, "sunet: patch_size %d unsupported (4)", m.patch);
  if (m.in_chans != 3) return fail(SUNET_E_SHAPE, "sunet: in_chans %d unsupported (3; grey inputs are repeated)", m.in_chans);
  if (m.img % 256) return fail(SUNET_E_SHAPE, "sunet: img_size %d must be a multiple of 256 (8x8 windows at 4 scales)", m.img);
  m.G = m.img / m.patch;
  const int E = m.E, G = m.G;
  // conv_first o patch_embed.proj -> 6x6 / stride 4 / pad 1 (exact linear fold)
  const float *w1, *b1, *w2, *b2;
  SUNET_TRY(P.get("conv_first.weight", static_cast<int64_t>(E) * 3 * 9, &w1));
  SUNET_TRY(P.get("conv_first.bias", E, &b1));
  SUNET_TRY(P.get("patch_embed.proj.weight", static_cast<int64_t>(E) * E * 16, &w2));
  SUNET_TRY(P.get("patch_embed.proj.bias", E, &b2));
  SUNET_TRY(ar.alloc_t(&m.wfold, static_cast<size_t>(108) * E));
  SUNET_TRY(ar.alloc_t(&m.bfold, E));
  SUNET_TRY(fold_patch_embed(w1, b1, w2, b2, 3, E, m.wfold, m.bfold, s));
  SUNET_TRY(ar.alloc_t(&m.wpk, static_cast<size_t>(E) * PATCH_EMBED_WPK_PITCH));
  SUNET_TRY(pack_patch_embed_f16(m.wfold, m.wpk, E, s));
  if (P.has("patch_embed.norm.weight")) {
    SUNET_TRY(copy_vec(ar, P, "patch_embed.norm.weight", E, &m.pe_g, s));
    SUNET_TRY(copy_vec(ar, P, "patch_embed.norm.bias", E, &m.pe_b, s));
  } else {
    return fail(SUNET_E_ARG, "sunet: patch_norm=False is not supported by the fused patch embed");
  }
  for (int i = 0; i < 4; ++i) {
    const int dim = E << i, H = G >> i;
    m.enc[i].resize(m.depths[i]);
    for (int j = 0; j < m.depths[i]; ++j) {
      const std::string pre = "layers." + std::to_string(i) + ".blocks." + std::to_string(j) + ".";
      SUNET_TRY(m.enc[i][j].pack(ar, P, pre, dim, H, H, m.heads[i], (j % 2 == 0) ? 0 : 4, qk_scale, s));  // :423
    }
    if (i < 3) SUNET_TRY(m.merge[i].pack(ar, P, "layers." + std::to_string(i) + ".downsample.", dim, H, H, s));
  }
  SUNET_TRY(copy_vec(ar, P, "norm.weight", E * 8, &m.norm_g, s));
  SUNET_TRY(copy_vec(ar, P, "norm.bias", E * 8, &m.norm_b, s));
  SUNET_TRY(copy_vec(ar, P, "norm_up.weight", E, &m.normup_g, s));
  SUNET_TRY(copy_vec(ar, P, "norm_up.bias", E, &m.normup_b, s));
  SUNET_TRY(m.up0.pack(ar, P, "layers_up.0.", E * 8, 2, G >> 3, G >> 3, s));
  for (int inx = 1; inx < 4; ++inx) {
    const int i = 3 - inx;
    const int dim = E << i, H = G >> i;
    SUNET_TRY(pack_linear(ar, P, "concat_back_dim." + std::to_string(inx) + ".weight", "concat_back_dim." + std::to_string(inx) + ".bias",
                          dim, 2 * dim, &m.cat[inx], s));
    m.dec[inx].resize(m.depths[i]);
    for (int j = 0; j < m.depths[i]; ++j) {
      const std::string pre = "layers_up." + std::to_string(inx) + ".blocks." + std::to_string(j) + ".";
      SUNET_TRY(m.dec[inx][j].pack(ar, P, pre, dim, H, H, m.heads[i], (j % 2 == 0) ? 0 : 4, qk_scale, s));  // :493
    }
    if (inx < 3) SUNET_TRY(m.ups[inx].pack(ar, P, "layers_up." + std::to_string(inx) + ".upsample.", dim, 2, H, H, s));
  }
  SUNET_TRY(m.tail.pack(ar, P, "up.", "output.weight", E, m.out_chans, G, G, s));
  return 0;
}

// one chunk of Bc images; x NCHW fp32 -> out NCHW fp32
static int model_forward_chunk(const ModelPack& m, Ctx& c, const void* x, int in_chans, int Bc, void* out) {
  ScratchMark mk(c.sc);
  const int E = m.E, G = m.G;
  const int64_t S = static_cast<int64_t>(Bc) * G * G * E;  // elements of the stage-0 token stream
  __half *skip[3], *Xa, *Xb, *T;
  SUNET_TRY(c.sc.take_t(&skip[0], S));
  SUNET_TRY(c.sc.take_t(&skip[1], S / 2));
  SUNET_TRY(c.sc.take_t(&skip[2], S / 4));
  SUNET_TRY(c.sc.take_t(&Xa, S));
  SUNET_TRY(c.sc.take_t(&Xb, S));
  SUNET_TRY(c.sc.take_t(&T, S));
  RUN(c, K_PATCH_EMBED, 2.0 * 108 * S, (c.in_fmt == IMG_U8_NHWC ? 1.0 : 4.0) * Bc * in_chans * m.img * m.img + 2.0 * S,
      patch_embed_fused(x, c.in_fmt, in_chans, Bc, m.img, m.img, m.wfold, m.wpk, m.bfold, m.pe_g, m.pe_b, E, skip[0], c.stream));  // :749, :708
  // encoder + bottleneck (:714-716).  x_downsample[i] = input of stage i stays untouched in skip[i].
  const __half* cur = skip[0];
  for (int i = 0; i < 4; ++i) {
    for (size_t j = 0; j < m.enc[i].size(); ++j) {
      SUNET_TRY(m.enc[i][j].forward(c, cur, Xa, Bc));
      cur = Xa;
    }
    if (i < 3) {
      __half* dst = i < 2 ? skip[i + 1] : Xb;   // the input of stage 3 is not a skip (only x_downsample[0..2] are read, :728)
      SUNET_TRY(m.merge[i].forward(c, cur, dst, Bc));
      cur = dst;
    }
  }
  const int64_t M3 = static_cast<int64_t>(Bc) * (G >> 3) * (G >> 3);
  RUN(c, K_LAYERNORM, 0.0, 32.0 * M3 * E, layernorm_f16(cur, 8 * E, T, 8 * E, m.norm_g, m.norm_b, M3, 8 * E, c.stream));  // :718
  SUNET_TRY(m.up0.forward(c, T, Xb, 0, Bc));                                             // :726
  for (int inx = 1; inx < 4; ++inx) {
    const int i = 3 - inx;
    const int dim = E << i, H = G >> i;
    const int64_t M = static_cast<int64_t>(Bc) * H * H;
    // cat([x, skip], -1) -> Linear(2C -> C) as two accumulating K-segments (:728-729)
    SUNET_TRY(run_linear(c, m.cat[inx], Xb, dim, M, Xa, dim, ACT_NONE, nullptr, nullptr, 0, 0, skip[i], dim, dim));
    for (size_t j = 0; j < m.dec[inx].size(); ++j) SUNET_TRY(m.dec[inx][j].forward(c, Xa, Xa, Bc));
    if (inx < 3) SUNET_TRY(m.ups[inx].forward(c, Xa, Xb, 0, Bc));
  }
  const int64_t M0 = static_cast<int64_t>(Bc) * G * G;
  RUN(c, K_LAYERNORM, 0.0, 4.0 * M0 * E, layernorm_f16(Xa, E, T, E, m.normup_g, m.normup_b, M0, E, c.stream));            // :732
  SUNET_TRY(m.tail.forward(c, T, out, Bc));                                              // :742-753
  return 0;
}

// x / out: c.in_fmt / c.out_fmt images (fp32 NCHW or 8-bit NHWC); c.ev (optional) describes the whole batch.
static int model_forward(const ModelPack& m, Ctx& c, const void* x, int in_chans, int batch, int max_chunk, void* out) {
  if (in_chans != 1 && in_chans != 3) return fail(SUNET_E_SHAPE, "sunet: input has %d channels (1 or 3)", in_chans);
  if (batch <= 0) return fail(SUNET_E_SHAPE, "sunet: batch %d", batch);
  if (max_chunk <= 0) max_chunk = 64;
  const int64_t plane = static_cast<int64_t>(m.img) * m.img;
  const int64_t in_img = in_chans * plane * (c.in_fmt == IMG_U8_NHWC ? 1 : 4);        // bytes per image
  const int64_t out_img = m.out_chans * plane * (c.out_fmt == IMG_U8_NHWC ? 1 : 4);
  const EvalEpilogue* whole = c.ev;
  int rc = 0;
  for (int b0 = 0; b0 < batch && !rc; b0 += max_chunk) {
    const int bc = batch - b0 < max_chunk ? batch - b0 : max_chunk;
    EvalEpilogue part;
    if (whole) {
      part = *whole;
      part.target += b0 * whole->target_chans * plane;
      if (part.weight) part.weight += b0 * plane;
      if (part.prob) part.prob += b0 * m.out_chans * plane;
      c.ev = &part;
    }
    rc = model_forward_chunk(m, c, x ? static_cast<const uint8_t*>(x) + b0 * in_img : nullptr, in_chans, bc,
                             out ? static_cast<uint8_t*>(out) + b0 * out_img : nullptr);
  }
  c.ev = whole;
  return rc;
}

// ------------------------------------------------------------------------------------------------ per-module fp32 wrappers
struct AsyncScratch {  // stream-ordered scratch for the per-module entry points
  void* p = nullptr;
  cudaStream_t s;
  explicit AsyncScratch(cudaStream_t st) : s(st) {}
  ~AsyncScratch() { if (p) cudaFreeAsync(p, s); }
  int alloc(size_t bytes) { SUNET_CUDA(cudaMallocAsync(&p, bytes, s)); return 0; }
};

template <typename F>
static int with_scratch(cudaStream_t stream, F&& body) {
  Ctx dry;
  dry.sc.dry = true;
  dry.stream = stream;
  SUNET_TRY(body(dry));
  AsyncScratch as(stream);
  SUNET_TRY(as.alloc(dry.sc.peak + 256));
  Ctx c;
  c.stream = stream;
  c.sc.base = static_cast<uint8_t*>(as.p);
  c.sc.cap = dry.sc.peak + 256;
  return body(c);
}

}  // namespace sunet

// ================================================================================================ C ABI
using namespace sunet;

extern "C" {

int sunet_abi_version(void) { return 1; }
const char* sunet_last_error(void) { return last_error_buf(); }

int sunet_prepack(const char* kind, const int64_t* ia, int nia, const double* fa, int nfa, const char* const* names,
                  const void* const* ptrs, const int64_t* numels, int n_params, void* stream, sunet_handle_t* out) {
  if (!kind || !out) return fail(SUNET_E_ARG, "prepack: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Params P;
  for (int i = 0; i < n_params; ++i) {
    if ((reinterpret_cast<uintptr_t>(ptrs[i]) & 3) != 0) return fail(SUNET_E_ALIGN, "parameter '%s' is not 4-byte aligned", names[i]);
    P.m[names[i]] = {static_cast<const float*>(ptrs[i]), numels[i]};
  }
  const std::string k = kind;
  const double f0 = nfa > 0 ? fa[0] : 0.0;
  std::unique_ptr<Handle> h;
  int rc = 0;
  auto need = [&](int n) { return nia >= n ? 0 : fail(SUNET_E_ARG, "prepack(%s): %d integer arguments expected, got %d", kind, n, nia); };
  if (k == "swin_block") {
    SUNET_TRY(need(5));
    auto* b = new BlockHandle(); h.reset(b);
    rc = b->p.pack(b->arena, P, "", (int)ia[0], (int)ia[1], (int)ia[2], (int)ia[3], (int)ia[4], f0, s);
  } else if (k == "window_attention") {
    SUNET_TRY(need(2));
    auto* a = new AttnHandle(); h.reset(a);
    rc = a->p.pack(a->arena, P, "", (int)ia[0], (int)ia[1], f0, s);
  } else if (k == "mlp") {
    SUNET_TRY(need(3));
    auto* m = new MlpHandle(); h.reset(m);
    rc = m->p.pack(m->arena, P, "", (int)ia[0], (int)ia[1], (int)ia[2], s);
  } else if (k == "patch_merging") {
    SUNET_TRY(need(3));
    auto* m = new MergeHandle(); h.reset(m);
    rc = m->p.pack(m->arena, P, "", (int)ia[0], (int)ia[1], (int)ia[2], s);
  } else if (k == "upsample") {
    SUNET_TRY(need(4));
    auto* u = new UpHandle(); h.reset(u);
    rc = u->p.pack(u->arena, P, "", (int)ia[0], (int)ia[1], (int)ia[2], (int)ia[3], s);
  } else if (k == "patch_embed") {
    SUNET_TRY(need(4));
    auto* e = new PatchEmbedHandle(); h.reset(e);
    PatchEmbedPack& p = e->p;
    p.cin = (int)ia[0]; p.E = (int)ia[1]; p.P = (int)ia[2]; p.has_norm = (int)ia[3];
    if ((p.cin * p.P * p.P) % 16 || p.E % 16) rc = fail(SUNET_E_SHAPE, "patch_embed: in_chans*patch^2 and embed_dim must be multiples of 16");
    if (!rc) rc = pack_linear(e->arena, P, "proj.weight", "proj.bias", p.E, p.cin * p.P * p.P, &p.proj, s);
    if (!rc && p.has_norm) {
      rc = copy_vec(e->arena, P, "norm.weight", p.E, &p.g, s);
      if (!rc) rc = copy_vec(e->arena, P, "norm.bias", p.E, &p.b, s);
    }
  } else if (k == "sunet") {
    auto* m = new ModelHandle(); h.reset(m);
    rc = pack_model(m, P, ia, nia, f0, s);
  } else {
    return fail(SUNET_E_ARG, "prepack: unknown kind '%s'", kind);
  }
  if (rc) return rc;
  h->kind = k;
  SUNET_CUDA(cudaStreamSynchronize(s));  // the caller may free / mutate its fp32 parameters after this returns
  *out = reinterpret_cast<sunet_handle_t>(h.release());
  return 0;
}

int sunet_destroy(sunet_handle_t h) {
  delete reinterpret_cast<Handle*>(h);
  return 0;
}

#define GET_HANDLE(T, var, h, kindstr)                                                     \
  Handle* _base = reinterpret_cast<Handle*>(h);                                            \
  if (!_base || _base->kind != kindstr) return fail(SUNET_E_ARG, "handle is not a %s", kindstr); \
  T* var = static_cast<T*>(_base);

int sunet_swin_block_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream) {
  GET_HANDLE(BlockHandle, bh, h, "swin_block");
  const BlockPack& p = bh->p;
  const int64_t n = static_cast<int64_t>(batch) * p.H * p.W * p.dim;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half *xi, *xo;
    SUNET_TRY(c.sc.take_t(&xi, n));
    SUNET_TRY(c.sc.take_t(&xo, n));
    RUN(c, K_CAST, 0.0, 6.0 * (n), cast_f32_to_f16(x, xi, n, c.stream));
    SUNET_TRY(p.forward(c, xi, xo, batch));
    RUN(c, K_CAST, 0.0, 6.0 * (n), cast_f16_to_f32(xo, out, n, c.stream));
    return 0;
  });
}

size_t sunet_swin_block_f16_workspace_bytes(sunet_handle_t h, int batch) {
  Handle* base = reinterpret_cast<Handle*>(h);
  if (!base || base->kind != "swin_block") { fail(SUNET_E_ARG, "handle is not a swin_block"); return 0; }
  const BlockPack& p = static_cast<BlockHandle*>(base)->p;
  Ctx dry;
  dry.sc.dry = true;
  if (p.forward(dry, nullptr, nullptr, batch)) return 0;
  return dry.sc.peak + 256;
}

int sunet_swin_block_f16(sunet_handle_t h, const void* x, int batch, int part, void* out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  GET_HANDLE(BlockHandle, bh, h, "swin_block");
  const BlockPack& p = bh->p;
  if (!x || !out || !workspace) return fail(SUNET_E_ARG, "sunet_swin_block_f16: null pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SUNET_E_ALIGN, "sunet_swin_block_f16: workspace must be 256-byte aligned");
  if (batch <= 0) return fail(SUNET_E_SHAPE, "sunet_swin_block_f16: batch %d", batch);
  if (part != 0 && part != 1) return fail(SUNET_E_ARG, "sunet_swin_block_f16: part %d (0 whole block, 1 attention part)", part);
  Ctx c;
  c.stream = static_cast<cudaStream_t>(stream);
  c.sc.base = static_cast<uint8_t*>(workspace);
  c.sc.cap = workspace_bytes;
  const __half* xi = static_cast<const __half*>(x);
  __half* xo = static_cast<__half*>(out);
  if (part == 0) return p.forward(c, xi, xo, batch);
  const int64_t M = static_cast<int64_t>(batch) * p.H * p.W;
  __half *T, *QKV;
  SUNET_TRY(c.sc.take_t(&T, p.use_af ? 0 : M * p.dim));
  SUNET_TRY(c.sc.take_t(&QKV, p.use_af ? 0 : M * 3 * p.dim));
  return p.attention_part(c, xi, xo, T, QKV, batch);
}

int sunet_window_attention_fwd(sunet_handle_t h, const float* x, int64_t num_windows, const float* mask, int mask_nw, float* out,
                               void* stream) {
  GET_HANDLE(AttnHandle, ah, h, "window_attention");
  const AttnPack& p = ah->p;
  if (mask && (mask_nw <= 0 || num_windows % mask_nw)) return fail(SUNET_E_SHAPE, "window attention: %lld windows not divisible by mask nW=%d", (long long)num_windows, mask_nw);
  const int64_t M = num_windows * 64, n = M * p.dim;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half *xi, *QKV, *O, *Y;
    SUNET_TRY(c.sc.take_t(&xi, n));
    SUNET_TRY(c.sc.take_t(&QKV, 3 * n));
    SUNET_TRY(c.sc.take_t(&O, n));
    SUNET_TRY(c.sc.take_t(&Y, n));
    RUN(c, K_CAST, 0.0, 6.0 * (n), cast_f32_to_f16(x, xi, n, c.stream));
    SUNET_TRY(run_linear(c, p.qkv, xi, p.dim, M, QKV, 3 * p.dim));
    AttnCoreArgs a;
    a.qkv = QKV; a.ld = 3 * p.dim; a.out = O; a.ldo = p.dim; a.B = 1; a.H = 8; a.W = 8; a.C = p.dim; a.heads = p.heads;
    a.bias_table = p.table; a.windowed_input = 1; a.num_windows = num_windows;
    a.mask_mode = mask ? 2 : 0; a.mask = mask; a.mask_nw = mask_nw;
    RUN(c, K_ATTN, 256.0 * M * p.dim, 8.0 * M * p.dim, attn_core_launch(a, c.stream));
    SUNET_TRY(run_linear(c, p.proj, O, p.dim, M, Y, p.dim));
    RUN(c, K_CAST, 0.0, 6.0 * (n), cast_f16_to_f32(Y, out, n, c.stream));
    return 0;
  });
}

int sunet_mlp_fwd(sunet_handle_t h, const float* x, int64_t rows, float* out, void* stream) {
  GET_HANDLE(MlpHandle, mh, h, "mlp");
  const MlpPack& p = mh->p;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half *xi, *Hd, *Y;
    SUNET_TRY(c.sc.take_t(&xi, rows * p.cin));
    SUNET_TRY(c.sc.take_t(&Hd, rows * p.hid));
    SUNET_TRY(c.sc.take_t(&Y, rows * p.cout));
    RUN(c, K_CAST, 0.0, 6.0 * (rows * p.cin), cast_f32_to_f16(x, xi, rows * p.cin, c.stream));
    SUNET_TRY(run_linear(c, p.fc1, xi, p.cin, rows, Hd, p.hid, ACT_GELU));
    SUNET_TRY(run_linear(c, p.fc2, Hd, p.hid, rows, Y, p.cout));
    RUN(c, K_CAST, 0.0, 6.0 * (rows * p.cout), cast_f16_to_f32(Y, out, rows * p.cout, c.stream));
    return 0;
  });
}

int sunet_patch_merging_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream) {
  GET_HANDLE(MergeHandle, mh, h, "patch_merging");
  const MergePack& p = mh->p;
  const int64_t n_in = static_cast<int64_t>(batch) * p.H * p.W * p.dim, n_out = n_in / 2;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half *xi, *Y;
    SUNET_TRY(c.sc.take_t(&xi, n_in));
    SUNET_TRY(c.sc.take_t(&Y, n_out));
    RUN(c, K_CAST, 0.0, 6.0 * (n_in), cast_f32_to_f16(x, xi, n_in, c.stream));
    SUNET_TRY(p.forward(c, xi, Y, batch));
    RUN(c, K_CAST, 0.0, 6.0 * (n_out), cast_f16_to_f32(Y, out, n_out, c.stream));
    return 0;
  });
}

int sunet_upsample_fwd(sunet_handle_t h, const float* x, int batch, float* out, void* stream) {
  GET_HANDLE(UpHandle, uh, h, "upsample");
  const UpPack& p = uh->p;
  const int64_t n_in = static_cast<int64_t>(batch) * p.H * p.W * p.C;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half* xi;
    SUNET_TRY(c.sc.take_t(&xi, n_in));
    RUN(c, K_CAST, 0.0, 6.0 * (n_in), cast_f32_to_f16(x, xi, n_in, c.stream));
    SUNET_TRY(p.forward(c, xi, out, 1, batch));  // raster NHWC == (B, 4L, C/2) for r=2 and (B, 4H, 4W, C) for r=4
    return 0;
  });
}

int sunet_patch_embed_fwd(sunet_handle_t h, const float* x, int batch, int himg, int wimg, float* out, void* stream) {
  GET_HANDLE(PatchEmbedHandle, eh, h, "patch_embed");
  const PatchEmbedPack& p = eh->p;
  if (himg % p.P || wimg % p.P) return fail(SUNET_E_SHAPE, "patch_embed: image %dx%d not divisible by patch %d", himg, wimg, p.P);
  const int64_t M = static_cast<int64_t>(batch) * (himg / p.P) * (wimg / p.P);
  const int K = p.cin * p.P * p.P;
  return with_scratch(static_cast<cudaStream_t>(stream), [&](Ctx& c) -> int {
    __half *A, *Y, *Yn;
    SUNET_TRY(c.sc.take_t(&A, M * K));
    SUNET_TRY(c.sc.take_t(&Y, M * p.E));
    SUNET_TRY(c.sc.take_t(&Yn, M * p.E));
    RUN(c, K_IM2COL, 0.0, 6.0 * M * K, im2col_patch(x, batch, p.cin, himg, wimg, p.P, A, c.stream));
    SUNET_TRY(run_linear(c, p.proj, A, K, M, Y, p.E));
    if (p.has_norm) {
      RUN(c, K_LAYERNORM, 0.0, 4.0 * M * p.E, layernorm_f16(Y, p.E, Yn, p.E, p.g, p.b, M, p.E, c.stream));
      RUN(c, K_CAST, 0.0, 6.0 * (M * p.E), cast_f16_to_f32(Yn, out, M * p.E, c.stream));
    } else {
      RUN(c, K_CAST, 0.0, 6.0 * (M * p.E), cast_f16_to_f32(Y, out, M * p.E, c.stream));
    }
    return 0;
  });
}

size_t sunet_workspace_bytes(sunet_handle_t h, int batch, int max_chunk) {
  Handle* base = reinterpret_cast<Handle*>(h);
  if (!base || base->kind != "sunet") { fail(SUNET_E_ARG, "handle is not a sunet model"); return 0; }
  const ModelPack& m = static_cast<ModelHandle*>(base)->p;
  Ctx dry;
  dry.sc.dry = true;
  if (model_forward(m, dry, nullptr, 3, batch, max_chunk, nullptr)) return 0;
  return dry.sc.peak + 256;
}

int64_t sunet_forward_launches(sunet_handle_t h, int batch, int max_chunk) {
  Handle* base = reinterpret_cast<Handle*>(h);
  if (!base || base->kind != "sunet") { fail(SUNET_E_ARG, "handle is not a sunet model"); return -1; }
  const ModelPack& m = static_cast<ModelHandle*>(base)->p;
  Ctx dry;
  dry.sc.dry = true;
  if (model_forward(m, dry, nullptr, 3, batch, max_chunk, nullptr)) return -1;
  return dry.launches;
}

int sunet_forward(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, float* out, void* workspace,
                  size_t workspace_bytes, void* stream) {
  GET_HANDLE(ModelHandle, mh, h, "sunet");
  if (!x || !out || !workspace) return fail(SUNET_E_ARG, "sunet_forward: null pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SUNET_E_ALIGN, "sunet_forward: workspace must be 256-byte aligned");
  Ctx c;
  c.stream = static_cast<cudaStream_t>(stream);
  c.sc.base = static_cast<uint8_t*>(workspace);
  c.sc.cap = workspace_bytes;
  return model_forward(mh->p, c, x, in_chans, batch, max_chunk, out);
}

int sunet_forward_u8(sunet_handle_t h, const uint8_t* x, int in_chans, int batch, int max_chunk, uint8_t* out, void* workspace,
                     size_t workspace_bytes, void* stream) {
  GET_HANDLE(ModelHandle, mh, h, "sunet");
  if (!x || !out || !workspace) return fail(SUNET_E_ARG, "sunet_forward_u8: null pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SUNET_E_ALIGN, "sunet_forward_u8: workspace must be 256-byte aligned");
  Ctx c;
  c.stream = static_cast<cudaStream_t>(stream);
  c.sc.base = static_cast<uint8_t*>(workspace);
  c.sc.cap = workspace_bytes;
  c.in_fmt = c.out_fmt = IMG_U8_NHWC;
  return model_forward(mh->p, c, x, in_chans, batch, max_chunk, out);
}

int sunet_forward_eval(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, const float* target, int target_chans,
                       const float* weight, float eps, float* logits, float* prob, double* sums, void* workspace, size_t workspace_bytes,
                       void* stream) {
  GET_HANDLE(ModelHandle, mh, h, "sunet");
  if (!x || !target || !logits || !sums || !workspace) return fail(SUNET_E_ARG, "sunet_forward_eval: null pointer");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SUNET_E_ALIGN, "sunet_forward_eval: workspace must be 256-byte aligned");
  if (target_chans != mh->p.out_chans && !(target_chans == 3 && mh->p.out_chans == 1))
    return fail(SUNET_E_SHAPE, "sunet_forward_eval: target has %d channels, the model outputs %d", target_chans, mh->p.out_chans);
  Ctx c;
  c.stream = static_cast<cudaStream_t>(stream);
  c.sc.base = static_cast<uint8_t*>(workspace);
  c.sc.cap = workspace_bytes;
  EvalEpilogue ev;
  ev.target = target; ev.target_chans = target_chans; ev.weight = weight; ev.prob = prob; ev.sums = sums; ev.eps = eps;
  c.ev = &ev;
  SUNET_CUDA(cudaMemsetAsync(sums, 0, 5 * sizeof(double), c.stream));
  return model_forward(mh->p, c, x, in_chans, batch, max_chunk, logits);
}

int sunet_forward_profile(sunet_handle_t h, const float* x, int in_chans, int batch, int max_chunk, float* out, void* workspace,
                          size_t workspace_bytes, void* stream, sunet_prof_rec* recs, int max_recs, int* n_recs) {
  GET_HANDLE(ModelHandle, mh, h, "sunet");
  if (!x || !out || !workspace || !recs || !n_recs) return fail(SUNET_E_ARG, "sunet_forward_profile: null pointer");
  std::vector<ProfRec> prof;
  Ctx c;
  c.stream = static_cast<cudaStream_t>(stream);
  c.sc.base = static_cast<uint8_t*>(workspace);
  c.sc.cap = workspace_bytes;
  c.prof = &prof;
  int rc = model_forward(mh->p, c, x, in_chans, batch, max_chunk, out);
  cudaError_t e = cudaStreamSynchronize(c.stream);
  if (!rc && e != cudaSuccess) rc = fail((int)e, "sunet_forward_profile: %s", cudaGetErrorString(e));
  int n = 0;
  for (ProfRec& r : prof) {
    if (!rc && n < max_recs) {
      float ms = 0.f;
      if (r.e1 && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        recs[n].kind = r.kind; recs[n].ms = ms; recs[n].flops = r.flops; recs[n].bytes = r.bytes;
        ++n;
      }
    }
    if (r.e0) cudaEventDestroy(r.e0);
    if (r.e1) cudaEventDestroy(r.e1);
  }
  *n_recs = n;
  return rc;
}

int sunet_selftest_umma(void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int N = 96;
  std::vector<__half> hA(128 * 64), hB(N * 64);
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = __float2half(((int)(i * 37 % 101) - 50) / 64.f);
  for (size_t i = 0; i < hB.size(); ++i) hB[i] = __float2half(((int)(i * 53 % 89) - 44) / 64.f);
  __half *dA, *dB;
  float* dD;
  SUNET_CUDA(cudaMalloc(&dA, hA.size() * 2));
  SUNET_CUDA(cudaMalloc(&dB, hB.size() * 2));
  SUNET_CUDA(cudaMalloc(&dD, 128 * N * 4));
  SUNET_CUDA(cudaMemcpyAsync(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice, s));
  SUNET_CUDA(cudaMemcpyAsync(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice, s));
  int rc = umma_selftest(dA, dB, dD, N, s);
  std::vector<float> D(128 * N);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail((int)e, "selftest: %s", cudaGetErrorString(e));
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  if (rc) return rc;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      float ref = 0.f;
      for (int k = 0; k < 64; ++k) ref += __half2float(hA[i * 64 + k]) * __half2float(hB[j * 64 + k]);
      if (fabsf(ref - D[i * N + j]) > 1e-3f) return fail(SUNET_E_STATE, "umma selftest mismatch at (%d,%d): %f vs %f", i, j, D[i * N + j], ref);
    }
  return 0;
}

int sunet_gemm_f16(const void* A, const void* W, const float* bias, void* C, int64_t M, int N, int K, int act, void* stream) {
  GemmArgs a;
  a.A0 = static_cast<const __half*>(A); a.lda0 = K; a.K0 = K;
  a.W = static_cast<const __half*>(W); a.ldw = K; a.M = M; a.N = N; a.bias = bias; a.act = act;
  a.C = C; a.ldc = N;
  if (act == ACT_PRELU) return fail(SUNET_E_ARG, "sunet_gemm_f16: PReLU not exposed here");
  return gemm_run(a, static_cast<cudaStream_t>(stream));
}

int sunet_layernorm_f16(const void* x, int64_t rows, int C, const float* gamma, const float* beta, void* out, void* stream) {
  if (!x || !out || !gamma || !beta) return fail(SUNET_E_ARG, "sunet_layernorm_f16: null pointer");
  if (rows <= 0 || C <= 0 || C % 8) return fail(SUNET_E_SHAPE, "sunet_layernorm_f16: rows %lld, C %d (C must be a multiple of 8)", (long long)rows, C);
  return layernorm_f16(static_cast<const __half*>(x), C, static_cast<__half*>(out), C, gamma, beta, rows, C, static_cast<cudaStream_t>(stream));
}

int sunet_concat_linear_f16(const void* x, const void* skip, int64_t rows, int C, const float* w, const float* b, void* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!x || !skip || !w || !out) return fail(SUNET_E_ARG, "sunet_concat_linear_f16: null pointer");
  if (rows <= 0 || C <= 0 || C % 16) return fail(SUNET_E_SHAPE, "sunet_concat_linear_f16: rows %lld, C %d (C must be a multiple of 16)", (long long)rows, C);
  Arena ar;
  Linear L;
  L.N = C; L.K = 2 * C;
  SUNET_TRY(ar.alloc_t(&L.w, static_cast<size_t>(C) * 2 * C));
  SUNET_TRY(pack_weight_f16(w, L.w, C, 2 * C, 0, 1.f, s));
  if (b) {
    SUNET_TRY(ar.alloc_t(&L.b, C));
    SUNET_CUDA(cudaMemcpyAsync(L.b, b, C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  Ctx c;
  c.stream = s;
  // cat([x, skip], -1) -> Linear(2C -> C) as two accumulating K segments: the concatenated tensor never exists (:728-729)
  int rc = run_linear(c, L, static_cast<const __half*>(x), C, rows, out, C, ACT_NONE, nullptr, nullptr, 0, 0, static_cast<const __half*>(skip), C, C);
  cudaError_t e = cudaStreamSynchronize(s);   // the pack buffers are freed when `ar` goes out of scope
  if (!rc && e != cudaSuccess) rc = fail((int)e, "sunet_concat_linear_f16: %s", cudaGetErrorString(e));
  return rc;
}

int sunet_ln_mlp_residual_f16(const void* x, int64_t rows, int C, const float* gamma, const float* beta, const float* w1,
                              const float* b1, const float* w2, const float* b2, void* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!mlp_fused_supported(C)) return fail(SUNET_E_SHAPE, "fused mlp: C=%d not instantiated", C);
  Arena ar;
  MlpFusedPack mf;
  SUNET_TRY(ar.alloc_t(&mf.w1g, static_cast<size_t>(4) * C * C));
  SUNET_TRY(ar.alloc_t(&mf.w2, static_cast<size_t>(4) * C * C));
  SUNET_TRY(ar.alloc_t(&mf.hconst, static_cast<size_t>(8) * C));
  SUNET_TRY(ar.alloc_t(&mf.b2, C));
  SUNET_TRY(mlp_fused_prepack(&mf, C, gamma, beta, w1, b1, w2, b2, s));
  int rc = mlp_fused_launch(mf, static_cast<const __half*>(x), static_cast<__half*>(out), rows, s);
  cudaError_t e = cudaStreamSynchronize(s);   // the pack buffers are freed when `ar` goes out of scope
  if (!rc && e != cudaSuccess) rc = fail((int)e, "sunet_ln_mlp_residual_f16: %s", cudaGetErrorString(e));
  return rc;
}

}  // extern "C"
