// mc_dit_*: the whole patched forward of a Wan text-to-video model in ONE C call (SURVEY §8b `mc_dit_forward`):
//   prologue -> { cache hit: head(x0 + residual) | cache miss: block stack, residual = x - x0, head(x) }
// i.e. MagCache4Wan2.1/magcache_generate.py:229-275, :293-305 with the skip decision (:277-292, mc_ctrl_step) taken by the caller.
// Host code only: it issues exactly the launch sequence of WanEngine (magcache_b200/wan.py: prologue / _block / head) through the
// same C entry points of this library, on buffers carved out of one caller-owned workspace — same kernels, same operands, same
// order, hence bit-identical outputs (tests/test_native_forward_gpu.py). Built for the plain case the north-star workload is:
// one sample, one timestep, one GPU, 16 output channels; i2v / VACE / token-sharded / per-token-timestep forwards are sequenced by
// the Python engine.
// `mc_dit_plan` writes the launch plan as text without launching anything (operands by name): the CPU suite compares it line by
// line with the sequence the Python engine issues (tests/test_native_plan_cpu.py).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace mc {
namespace {

constexpr int64_t kAlign = 1024;
inline int64_t align_up(int64_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Range {
  const char* base;
  int64_t bytes;
  std::string name;
};

// workspace buffers, in layout order
enum Buf { B_TOK, B_X0, B_XS, B_H, B_ATT, B_FFN, B_CQ, B_QKV, B_CKV, B_CTXH, B_CTX, B_EM, B_SIN, B_EH, B_E, B_E0, B_PREP, B_ATTN_WS, B_COUNT };
const char* const kBufNames[B_COUNT] = {"tok", "x0", "xs", "h", "att", "ffn", "cq", "qkv", "ckv", "ctx_h", "ctx", "em", "sin", "e_h", "e", "e0",
                                        "head_prep", "attn_ws"};

}  // namespace
}  // namespace mc

struct mc_dit {
  mc_dit_dims d;
  mc_dit_weights w;
  std::vector<mc_dit_block> blocks;
  // bound shape
  int32_t F = 0, Hp = 0, Wp = 0;
  int64_t n_tok = 0;
  int64_t off[mc::B_COUNT + 1] = {};
  char* ws = nullptr;
  int64_t ws_bytes = 0, head_ws_bytes = 0, attn_ws_bytes = 0;
  const float* rope = nullptr;
};

namespace mc {
namespace {

int32_t layout(const mc_dit* h, int32_t F, int32_t Hp, int32_t Wp, int64_t* off, int64_t* head_ws, int64_t* attn_ws) {
  const mc_dit_dims& d = h->d;
  const int64_t N = static_cast<int64_t>(F) * Hp * Wp, D = d.dim, L = d.text_len;
  MC_CHECK_ARG(N <= INT32_MAX, "mc_dit: %lld tokens", static_cast<long long>(N));
  int32_t rc = mc_head_workspace_bytes(d.dim, head_ws);
  if (rc) return rc;
  int64_t a = 0, b = 0;
  rc = mc_attn_workspace_bytes(static_cast<int32_t>(N), static_cast<int32_t>(N), d.num_heads, &a);
  if (rc) return rc;
  rc = mc_attn_workspace_bytes(static_cast<int32_t>(N), d.text_len, d.num_heads, &b);
  if (rc) return rc;
  *attn_ws = a > b ? a : b;
  const int64_t sz[B_COUNT] = {
      N * d.in_dim * 4 * 2,  // tok   bf16 [N, in_dim*4]
      N * D * 2,             // x0    bf16 [N, D]
      N * D * 4,             // xs    fp32 [N, D]
      N * D * 2,             // h
      N * D * 2,             // att
      N * static_cast<int64_t>(d.ffn_dim) * 2,  // ffn
      N * D * 2,             // cq
      N * 3 * D * 2,         // qkv
      L * 2 * D * 2,         // ckv
      L * D * 2,             // ctx_h
      L * D * 2,             // ctx
      6 * D * 4,             // em
      static_cast<int64_t>(d.freq_dim) * 4,  // sin
      D * 4,                 // e_h
      D * 4,                 // e
      6 * D * 4,             // e0
      *head_ws,              // head_prep
      *attn_ws,              // attn_ws
  };
  int64_t o = 0;
  for (int i = 0; i < B_COUNT; ++i) {
    off[i] = o;
    o += align_up(sz[i] > 0 ? sz[i] : 1);
  }
  off[B_COUNT] = o;
  return MC_OK;
}

// One forward, either launched or written out as text. Every operand of the plan is printed as `name+byte_offset`.
struct Runner {
  mc_dit* h;
  void* stream;
  std::string* plan;  // non-null: plan mode, nothing is launched
  std::vector<Range> ranges;
  int32_t rc = MC_OK;

  void reg(const void* p, int64_t bytes, std::string name) { ranges.push_back({static_cast<const char*>(p), bytes, std::move(name)}); }
  std::string nm(const void* p) const {
    if (p == nullptr) return "null";
    const char* c = static_cast<const char*>(p);
    for (const Range& r : ranges)
      if (c >= r.base && c < r.base + r.bytes) return r.name + "+" + std::to_string(static_cast<long long>(c - r.base));
    return "?";
  }
  void line(const std::string& s) { plan->append(s).push_back('\n'); }
  static std::string I(int64_t v) { return std::to_string(static_cast<long long>(v)); }

  char* buf(Buf b) const { return h->ws + h->off[b]; }

#define MC_DIT_STEP(text, call) \
  do {                          \
    if (rc) return;             \
    if (plan)                   \
      line(text);               \
    else                        \
      rc = (call);              \
  } while (0)

  void patchify(const float* latent, void* tok) {
    const mc_dit_dims& d = h->d;
    MC_DIT_STEP("patchify " + nm(latent) + " C=" + I(d.in_dim) + " F=" + I(h->F) + " H=" + I(2 * h->Hp) + " W=" + I(2 * h->Wp) + " -> " + nm(tok),
                mc_patchify(latent, d.in_dim, h->F, 2 * h->Hp, 2 * h->Wp, tok, stream));
  }
  void gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, const float* bias, int32_t epi, void* out,
            int64_t ldo, const float* gate) {
    MC_DIT_STEP("gemm A=" + nm(A) + " lda=" + I(lda) + " B=" + nm(B) + " ldb=" + I(ldb) + " M=" + I(M) + " N=" + I(N) + " K=" + I(K) + " bias=" +
                    nm(bias) + " epi=" + I(epi) + " out=" + nm(out) + " ldo=" + I(ldo) + " gate=" + nm(gate),
                mc_gemm_bf16(A, lda, B, ldb, static_cast<int32_t>(M), static_cast<int32_t>(N), static_cast<int32_t>(K), bias, epi, out, ldo, gate,
                             stream));
  }
  void sinusoid(const double* t, float* out) {
    MC_DIT_STEP("time_sinusoid " + nm(t) + " n=1 dim=" + I(h->d.freq_dim) + " -> " + nm(out), mc_time_sinusoid(t, 1, h->d.freq_dim, out, stream));
  }
  void linear_small(const float* x, int64_t K, const float* W, const float* b, int64_t N, int32_t act, float* y) {
    MC_DIT_STEP("linear_f32_small x=" + nm(x) + " M=1 K=" + I(K) + " W=" + nm(W) + " b=" + nm(b) + " N=" + I(N) + " act=" + I(act) + " -> " + nm(y),
                mc_linear_f32_small(x, 1, static_cast<int32_t>(K), W, b, static_cast<int32_t>(N), act, y, stream));
  }
  void head_prepare() {
    const mc_dit_weights& w = h->w;
    MC_DIT_STEP("head_prepare mod=" + nm(w.head_mod) + " e=" + nm(buf(B_E)) + " Wt=" + nm(w.head_wt) + " b=" + nm(w.head_b) + " cols=" + I(h->d.dim) +
                    " -> " + nm(buf(B_PREP)),
                mc_head_prepare(w.head_mod, reinterpret_cast<const float*>(buf(B_E)), w.head_wt, w.head_b, h->d.dim, buf(B_PREP), h->head_ws_bytes,
                                stream));
  }
  void add(const void* x, int32_t xd, const void* r, int32_t rd, void* out, int32_t od, int64_t n) {
    MC_DIT_STEP("add x=" + nm(x) + ":" + I(xd) + " r=" + nm(r) + ":" + I(rd) + " -> " + nm(out) + ":" + I(od) + " n=" + I(n),
                mc_cache_hit_add(x, xd, r, rd, out, od, n, stream));
  }
  void residual_sub(const void* xo, const void* xi, float* r, int64_t n) {
    MC_DIT_STEP("residual_sub x_out=" + nm(xo) + " x_in=" + nm(xi) + " -> " + nm(r) + " n=" + I(n),
                mc_residual_sub(xo, MC_F32, xi, MC_BF16, r, MC_F32, n, stream));
  }
  void cast(const void* src, int32_t sd, void* dst, int32_t dd, int64_t n) {
    MC_DIT_STEP("cast " + nm(src) + ":" + I(sd) + " -> " + nm(dst) + ":" + I(dd) + " n=" + I(n), mc_cast(src, sd, dst, dd, n, stream));
  }
  void ln(const void* x, int64_t rows, int32_t mode, const float* p0, const float* p1, int32_t si, int32_t hi, int32_t round_bf16, void* out) {
    MC_DIT_STEP("ln_modulate x=" + nm(x) + " rows=" + I(rows) + " cols=" + I(h->d.dim) + " mode=" + I(mode) + " p0=" + nm(p0) + " p1=" + nm(p1) +
                    " scale=" + I(si) + " shift=" + I(hi) + " round=" + I(round_bf16) + " -> " + nm(out),
                mc_ln_modulate(x, MC_F32, rows, h->d.dim, h->d.eps, mode, p0, p1, si, hi, round_bf16, out, MC_BF16, stream));
  }
  void rms(void* x, int64_t ld, int64_t rows, int32_t segs, const float* w, const float* cos_sin) {
    if (segs > 1)
      MC_DIT_STEP("rmsnorm_rope_segs x=" + nm(x) + " ld=" + I(ld) + " rows=" + I(rows) + " segs=" + I(segs) + " cols=" + I(h->d.dim) + " w=" + nm(w) +
                      " rope=" + nm(cos_sin),
                  mc_rmsnorm_rope_segs(x, ld, rows, segs, h->d.dim, w, h->d.eps, cos_sin, 128, stream));
    else
      MC_DIT_STEP("rmsnorm_rope x=" + nm(x) + " ld=" + I(ld) + " rows=" + I(rows) + " cols=" + I(h->d.dim) + " w=" + nm(w) + " rope=" + nm(cos_sin),
                  mc_rmsnorm_rope(x, ld, rows, h->d.dim, w, h->d.eps, cos_sin, 128, stream));
  }
  void attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t Lq, int64_t Lk) {
    const float scale = 1.0f / std::sqrt(128.0f);
    MC_DIT_STEP("attention q=" + nm(q) + " ldq=" + I(ldq) + " k=" + nm(k) + " ldk=" + I(ldk) + " v=" + nm(v) + " ldv=" + I(ldv) + " out=" + nm(out) +
                    " Lq=" + I(Lq) + " Lk=" + I(Lk) + " heads=" + I(h->d.num_heads),
                mc_attn_fwd_ex(q, ldq, k, ldk, v, ldv, out, h->d.dim, static_cast<int32_t>(Lq), static_cast<int32_t>(Lk), h->d.num_heads, scale,
                               h->attn_ws_bytes > 0 ? buf(B_ATTN_WS) : nullptr, h->attn_ws_bytes, 0, nullptr, nullptr, 0, stream));
  }
  void head(const void* x, int32_t xd, const float* r, float* out) {
    float* outs[1] = {out};
    MC_DIT_STEP("head x=" + nm(x) + ":" + I(xd) + " r=" + nm(r) + " rows=" + I(h->n_tok) + " cols=" + I(h->d.dim) + " prep=" + nm(buf(B_PREP)) + " -> " +
                    nm(out),
                mc_head_unpatchify_ex(x, xd, r, h->n_tok, 0, h->d.dim, h->F, h->Hp, h->Wp, h->d.out_dim, h->d.eps, outs, 1, buf(B_PREP),
                                      h->head_ws_bytes, 0, stream));
  }
#undef MC_DIT_STEP

  void run(const float* latent, const double* t_dev, const void* context, int32_t skip, float* residual, float* out) {
    const mc_dit_dims& d = h->d;
    const mc_dit_weights& w = h->w;
    const int64_t N = h->n_tok, D = d.dim, Fd = d.ffn_dim, L = d.text_len;
    float* em = reinterpret_cast<float*>(buf(B_EM));
    float* e = reinterpret_cast<float*>(buf(B_E));
    float* e0 = reinterpret_cast<float*>(buf(B_E0));
    // ---- prologue (magcache_generate.py:229-275; WanEngine.prologue)
    patchify(latent, buf(B_TOK));
    gemm(buf(B_TOK), d.in_dim * 4, w.patch_w, d.in_dim * 4, N, D, d.in_dim * 4, w.patch_b, MC_EPI_BIAS_BF16, buf(B_X0), D, nullptr);
    sinusoid(t_dev, reinterpret_cast<float*>(buf(B_SIN)));
    linear_small(reinterpret_cast<float*>(buf(B_SIN)), d.freq_dim, w.time_w1, w.time_b1, D, 2, reinterpret_cast<float*>(buf(B_EH)));
    linear_small(reinterpret_cast<float*>(buf(B_EH)), D, w.time_w2, w.time_b2, D, 0, e);
    linear_small(e, D, w.tproj_w, w.tproj_b, 6 * D, 1, e0);
    head_prepare();
    if (skip) {  // :295 + :304-305 — the hit sum is formed inside the head kernel; the text embedding feeds only the blocks
      head(buf(B_X0), MC_BF16, residual, out);
      return;
    }
    gemm(context, d.text_dim, w.text_w1, d.text_dim, L, D, d.text_dim, w.text_b1, MC_EPI_BIAS_GELU_BF16, buf(B_CTXH), D, nullptr);
    gemm(buf(B_CTXH), D, w.text_w2, D, L, D, D, w.text_b2, MC_EPI_BIAS_BF16, buf(B_CTX), D, nullptr);
    // ---- block stack (:297-298; WanEngine.run_blocks / _block)
    cast(buf(B_X0), MC_BF16, buf(B_XS), MC_F32, N * D);
    char* qkv = buf(B_QKV);
    char* ckv = buf(B_CKV);
    for (int32_t li = 0; li < d.num_layers; ++li) {
      const mc_dit_block& b = h->blocks[li];
      add(b.mod, MC_F32, e0, MC_F32, em, MC_F32, 6 * D);
      ln(buf(B_XS), N, 0, em, nullptr, 1, 0, li == 0 ? 1 : 0, buf(B_H));
      gemm(buf(B_H), D, b.w_qkv, D, N, 3 * D, D, b.b_qkv, MC_EPI_BIAS_BF16, qkv, 3 * D, nullptr);
      rms(qkv, 3 * D, N, 2, b.nqk, h->rope);
      attention(qkv, 3 * D, qkv + D * 2, 3 * D, qkv + 2 * D * 2, 3 * D, buf(B_ATT), N, N);
      gemm(buf(B_ATT), D, b.w_o, D, N, D, D, b.b_o, MC_EPI_BIAS_GATE_RESID, buf(B_XS), D, em + 2 * D);
      ln(buf(B_XS), N, 1, b.n3_w, b.n3_b, 0, 0, 0, buf(B_H));
      gemm(buf(B_H), D, b.c_wq, D, N, D, D, b.c_bq, MC_EPI_BIAS_BF16, buf(B_CQ), D, nullptr);
      rms(buf(B_CQ), D, N, 1, b.c_nq, nullptr);
      gemm(buf(B_CTX), D, b.c_wkv, D, L, 2 * D, D, b.c_bkv, MC_EPI_BIAS_BF16, ckv, 2 * D, nullptr);
      rms(ckv, 2 * D, L, 1, b.c_nk, nullptr);
      attention(buf(B_CQ), D, ckv, 2 * D, ckv + D * 2, 2 * D, buf(B_ATT), N, L);
      gemm(buf(B_ATT), D, b.c_wo, D, N, D, D, b.c_bo, MC_EPI_BIAS_GATE_RESID, buf(B_XS), D, nullptr);
      ln(buf(B_XS), N, 0, em, nullptr, 4, 3, 0, buf(B_H));
      gemm(buf(B_H), D, b.w_f1, D, N, Fd, D, b.b_f1, MC_EPI_BIAS_GELU_BF16, buf(B_FFN), Fd, nullptr);
      gemm(buf(B_FFN), Fd, b.w_f2, Fd, N, D, Fd, b.b_f2, MC_EPI_BIAS_GATE_RESID, buf(B_XS), D, em + 5 * D);
    }
    residual_sub(buf(B_XS), buf(B_X0), residual, N * D);  // :299
    head(buf(B_XS), MC_F32, nullptr, out);                // :304-305
  }

  // names of everything a plan line can mention
  void register_names(const float* latent, const double* t_dev, const void* context, const float* residual, const float* out) {
    const mc_dit_dims& d = h->d;
    const mc_dit_weights& w = h->w;
    const int64_t D = d.dim, Fd = d.ffn_dim, N = h->n_tok;
    for (int i = 0; i < B_COUNT; ++i) reg(buf(static_cast<Buf>(i)), h->off[i + 1] - h->off[i], kBufNames[i]);
    reg(latent, N * d.in_dim * 4 * 4, "latent");
    reg(t_dev, 8, "t");
    reg(context, static_cast<int64_t>(d.text_len) * d.text_dim * 2, "context");
    reg(residual, N * D * 4, "residual");
    reg(out, N * d.out_dim * 4 * 4, "out");
    reg(h->rope, N * 128 * 4, "rope");
    reg(w.patch_w, D * d.in_dim * 4 * 2, "patch_w"), reg(w.patch_b, D * 4, "patch_b");
    reg(w.text_w1, D * d.text_dim * 2, "text_w1"), reg(w.text_b1, D * 4, "text_b1");
    reg(w.text_w2, D * D * 2, "text_w2"), reg(w.text_b2, D * 4, "text_b2");
    reg(w.time_w1, D * d.freq_dim * 4, "time_w1"), reg(w.time_b1, D * 4, "time_b1");
    reg(w.time_w2, D * D * 4, "time_w2"), reg(w.time_b2, D * 4, "time_b2");
    reg(w.tproj_w, 6 * D * D * 4, "tproj_w"), reg(w.tproj_b, 6 * D * 4, "tproj_b");
    reg(w.head_mod, 2 * D * 4, "head_mod"), reg(w.head_wt, D * 64 * 4, "head_wt"), reg(w.head_b, 64 * 4, "head_b");
    for (int32_t li = 0; li < d.num_layers; ++li) {
      const mc_dit_block& b = h->blocks[li];
      const std::string p = "blk" + std::to_string(li) + ".";
      reg(b.mod, 6 * D * 4, p + "mod");
      reg(b.w_qkv, 3 * D * D * 2, p + "w_qkv"), reg(b.b_qkv, 3 * D * 4, p + "b_qkv");
      reg(b.w_o, D * D * 2, p + "w_o"), reg(b.b_o, D * 4, p + "b_o");
      reg(b.nqk, 2 * D * 4, p + "nqk");
      reg(b.n3_w, D * 4, p + "n3_w"), reg(b.n3_b, D * 4, p + "n3_b");
      reg(b.c_wq, D * D * 2, p + "c_wq"), reg(b.c_bq, D * 4, p + "c_bq");
      reg(b.c_wkv, 2 * D * D * 2, p + "c_wkv"), reg(b.c_bkv, 2 * D * 4, p + "c_bkv");
      reg(b.c_wo, D * D * 2, p + "c_wo"), reg(b.c_bo, D * 4, p + "c_bo");
      reg(b.c_nq, D * 4, p + "c_nq"), reg(b.c_nk, D * 4, p + "c_nk");
      reg(b.w_f1, Fd * D * 2, p + "w_f1"), reg(b.b_f1, Fd * 4, p + "b_f1");
      reg(b.w_f2, D * Fd * 2, p + "w_f2"), reg(b.b_f2, D * 4, p + "b_f2");
    }
  }
};

int32_t check_bound(const mc_dit* h, const char* who) {
  MC_CHECK_ARG(h != nullptr, "%s: null handle", who);
  if (h->ws == nullptr) {
    set_error("%s: no workspace bound (mc_dit_bind)", who);
    return MC_ERR_STATE;
  }
  return MC_OK;
}

}  // namespace
}  // namespace mc

extern "C" mc_dit* mc_dit_create(const mc_dit_dims* dims, const mc_dit_weights* weights) {
  using namespace mc;
  if (dims == nullptr || weights == nullptr || weights->blocks == nullptr) {
    set_error("mc_dit_create: null argument");
    return nullptr;
  }
  const mc_dit_dims& d = *dims;
  if (d.dim < 128 || d.num_heads < 1 || d.dim != d.num_heads * 128 || d.ffn_dim < 8 || d.ffn_dim % 8 || d.num_layers < 1 || d.in_dim < 1 ||
      (d.in_dim * 4) % 8 || d.out_dim != 16 || d.freq_dim < 4 || d.freq_dim % 4 || d.text_dim < 8 || d.text_dim % 8 || d.text_len < 1 ||
      d.dim % 64 || !(d.eps > 0.0f)) {
    set_error("mc_dit_create: unsupported dims (dim=%d heads=%d [head_dim must be 128] ffn=%d layers=%d in=%d out=%d [must be 16] freq=%d text=%d x %d)",
              d.dim, d.num_heads, d.ffn_dim, d.num_layers, d.in_dim, d.out_dim, d.freq_dim, d.text_len, d.text_dim);
    return nullptr;
  }
  const mc_dit_weights& w = *weights;
  const void* top[] = {w.patch_w, w.patch_b, w.text_w1, w.text_b1, w.text_w2, w.text_b2, w.time_w1, w.time_b1,
                       w.time_w2, w.time_b2, w.tproj_w, w.tproj_b, w.head_mod, w.head_wt, w.head_b};
  for (const void* p : top)
    if (p == nullptr || !aligned16(p)) {
      set_error("mc_dit_create: a top-level weight pointer is null or not 16-byte aligned");
      return nullptr;
    }
  for (int32_t li = 0; li < d.num_layers; ++li) {
    const mc_dit_block& b = w.blocks[li];
    const void* ps[] = {b.mod, b.w_qkv, b.b_qkv, b.w_o, b.b_o, b.nqk, b.n3_w, b.n3_b, b.c_wq, b.c_bq,
                        b.c_wkv, b.c_bkv, b.c_wo, b.c_bo, b.c_nq, b.c_nk, b.w_f1, b.b_f1, b.w_f2, b.b_f2};
    for (const void* p : ps)
      if (p == nullptr || !aligned16(p)) {
        set_error("mc_dit_create: block %d has a null or misaligned weight pointer", li);
        return nullptr;
      }
  }
  mc_dit* h = new mc_dit();
  h->d = d;
  h->w = w;
  h->blocks.assign(w.blocks, w.blocks + d.num_layers);
  h->w.blocks = h->blocks.data();
  return h;
}

extern "C" void mc_dit_destroy(mc_dit* h) { delete h; }

extern "C" int32_t mc_dit_workspace_bytes(const mc_dit* h, int32_t F, int32_t Hp, int32_t Wp, int64_t* bytes_out) {
  MC_CHECK_ARG(h != nullptr && bytes_out != nullptr && F >= 1 && Hp >= 1 && Wp >= 1, "mc_dit_workspace_bytes: bad arguments");
  int64_t off[mc::B_COUNT + 1], hw = 0, aw = 0;
  const int32_t rc = mc::layout(h, F, Hp, Wp, off, &hw, &aw);
  if (rc) return rc;
  *bytes_out = off[mc::B_COUNT];
  return MC_OK;
}

extern "C" int32_t mc_dit_bind(mc_dit* h, int32_t F, int32_t Hp, int32_t Wp, void* workspace, int64_t workspace_bytes, const float* rope_cos_sin) {
  MC_CHECK_ARG(h != nullptr && F >= 1 && Hp >= 1 && Wp >= 1, "mc_dit_bind: bad arguments");
  MC_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & (mc::kAlign - 1)) == 0, "mc_dit_bind: workspace null or not 1024-byte aligned");
  MC_CHECK_ARG(rope_cos_sin != nullptr && (reinterpret_cast<uintptr_t>(rope_cos_sin) & 31u) == 0, "mc_dit_bind: rope table null or not 32-byte aligned");
  int64_t off[mc::B_COUNT + 1], hw = 0, aw = 0;
  const int32_t rc = mc::layout(h, F, Hp, Wp, off, &hw, &aw);
  if (rc) return rc;
  MC_CHECK_ARG(workspace_bytes >= off[mc::B_COUNT], "mc_dit_bind: workspace of %lld bytes, %lld needed (mc_dit_workspace_bytes)",
               static_cast<long long>(workspace_bytes), static_cast<long long>(off[mc::B_COUNT]));
  h->F = F, h->Hp = Hp, h->Wp = Wp, h->n_tok = static_cast<int64_t>(F) * Hp * Wp;
  std::memcpy(h->off, off, sizeof(off));
  h->ws = static_cast<char*>(workspace), h->ws_bytes = workspace_bytes, h->head_ws_bytes = hw, h->attn_ws_bytes = aw;
  h->rope = rope_cos_sin;
  return MC_OK;
}

extern "C" int32_t mc_dit_forward(mc_dit* h, const float* latent, const double* t_dev, const void* context_bf16, int32_t skip, float* residual,
                                  float* out, void* stream) {
  int32_t rc = mc::check_bound(h, "mc_dit_forward");
  if (rc) return rc;
  MC_CHECK_ARG(latent && t_dev && residual && out && (skip || context_bf16), "mc_dit_forward: null pointer");
  MC_CHECK_ARG(mc::aligned16(latent) && mc::aligned16(residual) && mc::aligned16(out) && (context_bf16 == nullptr || mc::aligned16(context_bf16)),
               "mc_dit_forward: latent / context / residual / out must be 16-byte aligned");
  mc::Runner r{h, stream, nullptr, {}};
  r.run(latent, t_dev, context_bf16, skip, residual, out);
  return r.rc;
}

extern "C" int32_t mc_dit_plan(mc_dit* h, int32_t skip, char* buf, int64_t buf_bytes, int64_t* needed) {
  int32_t rc = mc::check_bound(h, "mc_dit_plan");
  if (rc) return rc;
  MC_CHECK_ARG(needed != nullptr && (buf != nullptr || buf_bytes == 0), "mc_dit_plan: bad arguments");
  // the caller's tensors are not known here: stand-in addresses that cannot collide with real ranges, printed by name
  const char* fake = reinterpret_cast<const char*>(static_cast<uintptr_t>(16));
  const int64_t span = int64_t{1} << 40;
  std::string text;
  mc::Runner r{h, nullptr, &text, {}};
  const float* latent = reinterpret_cast<const float*>(fake);
  const double* t = reinterpret_cast<const double*>(fake + span);
  const void* ctx = fake + 2 * span;
  float* residual = reinterpret_cast<float*>(const_cast<char*>(fake + 3 * span));
  float* out = reinterpret_cast<float*>(const_cast<char*>(fake + 4 * span));
  r.register_names(latent, t, ctx, residual, out);
  r.run(latent, t, ctx, skip, residual, out);
  *needed = static_cast<int64_t>(text.size()) + 1;
  if (buf_bytes >= *needed) std::memcpy(buf, text.c_str(), text.size() + 1);
  return MC_OK;
}
