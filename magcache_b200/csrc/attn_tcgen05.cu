// Non-causal flash-attention forward on tcgen05/TMEM, head_dim 128 — the self-/cross-attention of the Wan DiT block
// (reference call chain MagCache4Wan2.1/magcache_generate.py:297-298 -> WanSelfAttention.forward ->
// flash_attention / SDPA, upstream wan/modules/{model,attention}.py).
//
// One CTA = one 128-row query tile of one head; KV is streamed in 64-row tiles. Two CTAs are co-resident per SM
// (256 TMEM columns and ~112 KB smem each), so one CTA's softmax overlaps the other's MMAs.
//   warps 0-3  softmax    : thread = query row (TMEM lane). tcgen05.ld S, online softmax in the exp2 domain with lazy
//                           rescaling of O (only when the running max grows by > 8), P -> bf16 -> swizzled smem
//   warp 4     TMA        : Q once; K tile (2 boxes) and V^T tile (1 box) per iteration into 2-stage rings
//   warp 5     MMA issuer : S_j = Q K_j^T (M128 N64 K128, double-buffered in TMEM), O += P_j V_j (M128 N128 K64)
// V is consumed transposed (V^T [heads*128, Lk], produced directly by the V-projection GEMM) so that both MMAs see
// K-major operands — the same smem/UMMA descriptor path the GEMM kernel uses.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {

constexpr int kBQ = 128, kBKV = 64, kHD = 128;
constexpr int kQBytes = kBQ * kHD * 2;     // 32 KB (two 64-column boxes of 16 KB)
constexpr int kKBytes = kBKV * kHD * 2;    // 16 KB (two boxes of 8 KB)
constexpr int kVBytes = kHD * kBKV * 2;    // 16 KB (one box: 128 d-rows x 64 kv)
constexpr int kPBytes = kBQ * kBKV * 2;    // 16 KB
constexpr int kKVStages = 2;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kKVStages * kKBytes;
constexpr int kOffP = kOffV + kKVStages * kVBytes;
constexpr int kOffBar = kOffP + kPBytes;  // 112 KB
constexpr int kAttnSmem = kOffBar + 256;
constexpr int kAttnThreads = 192;
constexpr int kTmemCols = 256;  // S0 [0,64) S1 [64,128) O [128,256)
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct AttnParams {
  int Lq, Lk, heads;
  int splits;      // > 1: blockIdx.z owns a contiguous range of KV tiles and writes a partial result (split-KV)
  float* part_o;   // [splits][Lq][heads*128] fp32, each split's output normalised by its own row sum
  float2* part_ml; // [splits][Lq][heads]     (row max in the scaled log2 domain, row sum)
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  int64_t ldo;
};

// P_IN_TMEM: the bf16 probabilities are written back into the (already consumed) S columns of TMEM with tcgen05.st and the
// PV MMA takes its A operand from TMEM — no smem round trip, no generic->async proxy fence, and P is double-buffered for free
// (it lives in S buffer j&1), which removes the write-P -> PV -> pv_done -> write-next-P serialisation of the smem variant.
// VARIANT 0: P through shared memory (kept for A/B measurements); VARIANT 1 (default): P in TMEM.
// Tried, validated and dropped in round 1 because they were slower (profiles/r01_attention_experiments.md; sources in the git
// history): speculative exponentials, cross-tile software pipelining, staggered CTA start, 256-row CTAs with 128-wide KV
// tiles, split-row softmax with 8 softmax warps, 128-wide KV tiles with a single-buffered S.
constexpr int kDefaultPoly = 0;  // see profiles/r01_attention_experiments.md ("exponentials on the FMA pipe")

template <int VARIANT, int POLY>
__global__ void __launch_bounds__(kAttnThreads, 2)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* s_free = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  constexpr bool P_IN_TMEM = VARIANT >= 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ;
  const int head = blockIdx.y;
  const int total_tiles = (p.Lk + kBKV - 1) / kBKV;
  const int per_split = (total_tiles + p.splits - 1) / p.splits;
  const int t0 = blockIdx.z * per_split;                       // first (global) KV tile of this CTA
  const int n_tiles = min(per_split, total_tiles - t0);        // local tile count (>= 1, guaranteed by the host)

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_fwd_kernel: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_vt);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&s_free[s], 128);
    }
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(pv_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 5) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, kQBytes);
      ptx::tma_load_2d(smem + kOffQ, &tmap_q, q_full, head * kHD, q0);
      ptx::tma_load_2d(smem + kOffQ + kQBytes / 2, &tmap_q, q_full, head * kHD + 64, q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[s], kKBytes);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes, &tmap_k, &k_full[s], head * kHD, (t0 + j) * kBKV);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes + kKBytes / 2, &tmap_k, &k_full[s], head * kHD + 64, (t0 + j) * kBKV);
        ptx::mbar_wait(&v_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], kVBytes);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes, &tmap_vt, &v_full[s], (t0 + j) * kBKV, head * kHD);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(kBQ, kBKV);  // 128 x 64
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(kBQ, kHD);   // 128 x 128
      const uint32_t q_addr = ptx::smem_u32(smem + kOffQ);
      const uint32_t p_addr = ptx::smem_u32(smem + kOffP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint32_t k_addr = ptx::smem_u32(smem + kOffK + s * kKBytes);
        const uint32_t tmem_s = tmem_base + s * kBKV;
#pragma unroll
        for (int kk = 0; kk < kHD / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_sw128_kmajor(q_addr + (kk >> 2) * (kQBytes / 2)) + 2 * (kk & 3);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(k_addr + (kk >> 2) * (kKBytes / 2)) + 2 * (kk & 3);
          ptx::umma_ss(tmem_s, da, db, idesc_s, kk != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&k_empty[s]);
        ptx::umma_commit(&s_full[s]);
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          const int t = j + 1;
          ptx::mbar_wait(&k_full[t & 1], (t >> 1) & 1);
          // S buffer t&1 was last used by tile t-2: its softmax must have drained it. With P in TMEM that is implied by
          // p_full(t-2) (waited before PV(t-2) was issued) and the in-order execution of PV(t-2) before this MMA.
          if (!P_IN_TMEM && t >= 2) ptx::mbar_wait(&s_free[t & 1], ((t - 2) >> 1) & 1);
          ptx::tc_fence_after();
          issue_s(t);
        }
        ptx::mbar_wait(p_full, j & 1);
        ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + kOffV + (j & 1) * kVBytes);
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          const uint64_t db = ptx::umma_desc_sw128_kmajor(v_addr) + 2 * kk;
          if (P_IN_TMEM) {
            ptx::umma_ts(tmem_o, tmem_base + (j & 1) * kBKV + kk * 8, db, idesc_o, (j | kk) != 0 ? 1u : 0u);
          } else {
            const uint64_t da = ptx::umma_desc_sw128_kmajor(p_addr) + 2 * kk;
            ptx::umma_ss(tmem_o, da, db, idesc_o, (j | kk) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&v_empty[j & 1]);
        ptx::umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------ softmax warps -----------------------------------------------
    const int r = warp * 32 + lane;  // row inside the Q tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint8_t* p_row = smem + kOffP + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      ptx::mbar_wait(&s_full[b], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sreg[2][32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_sel + b * kBKV, sreg[0]);
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_sel + b * kBKV + 32, sreg[1]);
      ptx::tmem_ld_wait();
      if (!P_IN_TMEM) {
        ptx::tc_fence_before();
        ptx::mbar_arrive(&s_free[b]);
      }

      const int valid = p.Lk - (t0 + j) * kBKV;  // columns >= valid are padding (only possible on the last tile)
      if (valid < kBKV) {                  // warp-uniform, taken at most once per CTA
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (h * 32 + c >= valid) sreg[h][c] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;  // two chains of 3-input max: 0.5 instruction per element
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          mx0 = ptx::max3(mx0, __uint_as_float(sreg[h][c]), __uint_as_float(sreg[h][c + 1]));
          mx1 = ptx::max3(mx1, __uint_as_float(sreg[h][c + 2]), __uint_as_float(sreg[h][c + 3]));
        }
      const float mx = fmaxf(mx0, mx1);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      bool waited = false;
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = m_new > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O must be quiescent: PV(j-1) complete
          ptx::mbar_wait(pv_done, (j - 1) & 1);
          ptx::tc_fence_after();
          waited = true;
          const float factor = need ? ptx::ex2_approx(m - m_new) : 1.0f;
          if (need) {
            l *= factor;
            m = m_new;
          }
#pragma unroll 1
          for (int c = 0; c < kHD / 32; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32b_x32(tmem_o + lane_sel + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            ptx::tmem_st_32x32b_x32(tmem_o + lane_sel + c * 32, o);
          }
          ptx::tmem_st_wait();
        }
      }
      // P = 2^(s*scale - m), row sum in fp32 before the bf16 rounding (as flash-attention does).
      // Packed fp32x2 FMA / ADD: one FFMA2 + two MUFU.EX2 + one FADD2 + one bf16x2 pack per pair of elements.
      const uint64_t scale2 = ptx::pack_f32x2(p.scale_log2, p.scale_log2);
      const uint64_t negm2 = ptx::pack_f32x2(-m, -m);
      uint64_t sum2a = 0ull, sum2b = 0ull;  // (+0.f, +0.f)
      uint32_t packed[32];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float a0, a1, b0, b1;
          ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[h][c]), __uint_as_float(sreg[h][c + 1])), scale2, negm2), a0, a1);
          ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[h][c + 2]), __uint_as_float(sreg[h][c + 3])), scale2, negm2), b0, b1);
          // POLY / 4 of the exponentials are evaluated on the FMA pipe (ptx::ex2_emul_pair) to take them off the MUFU, which
          // this loop otherwise keeps as busy as the tensor pipe (profiles/r01_ncu_attn.md): 1 = every other (b0, b1) pair,
          // 2 = every (b0, b1) pair, 3 = those plus every other (a0, a1) pair.
          if (POLY >= 3 && (c & 4)) {
            ptx::ex2_emul_pair(a0, a1);
          } else {
            a0 = ptx::ex2_approx(a0);
            a1 = ptx::ex2_approx(a1);
          }
          if (POLY >= 2 || (POLY == 1 && (c & 4))) {
            ptx::ex2_emul_pair(b0, b1);
          } else {
            b0 = ptx::ex2_approx(b0);
            b1 = ptx::ex2_approx(b1);
          }
          sum2a = ptx::add_f32x2(sum2a, ptx::pack_f32x2(a0, a1));
          sum2b = ptx::add_f32x2(sum2b, ptx::pack_f32x2(b0, b1));
          packed[h * 16 + (c >> 1)] = pack_bf16x2(a0, a1);
          packed[h * 16 + (c >> 1) + 1] = pack_bf16x2(b0, b1);
        }
      float s0, s1, s2, s3;
      ptx::unpack_f32x2(sum2a, s0, s1);
      ptx::unpack_f32x2(sum2b, s2, s3);
      const float psum = (s0 + s1) + (s2 + s3);
      l += psum;
      if (P_IN_TMEM) {
        // P_j overwrites the first 32 columns of S buffer b (64 bf16 per row = 32 packed words); its reader PV(j) is ordered
        // before S(j+2) by the tensor pipe.
        // An mbarrier parity wait can only tell the current phase from the one before it, so a waiter must never fall two
        // phases behind. pv_done completes one phase per tile; a softmax thread that skipped it would reach the epilogue while
        // PV(n-2) is still in flight, and its wait for PV(n-1) would then be satisfied by the parity of PV(n-3): O read early,
        // run-to-run different results (caught by the full-shape determinism test, tools/diag_determinism.py). Observing
        // PV(j-1) here every tile costs nothing measurable — it has normally completed long before — and keeps the phases aligned.
        if (j > 0 && !waited) ptx::mbar_wait(pv_done, (j - 1) & 1);
        ptx::tmem_st_32x32b_x32(tmem_base + lane_sel + b * kBKV, packed);
        ptx::tmem_st_wait();
      } else {
        if (j > 0 && !waited) {
          ptx::mbar_wait(pv_done, (j - 1) & 1);  // PV(j-1) has finished reading the single smem P buffer
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 8 x 16-byte chunks per 128-byte row, XOR-swizzled with (row % 8)
          uint4 w = make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]);
          *reinterpret_cast<uint4*>(p_row + ((c ^ sw) << 4)) = w;
        }
        ptx::fence_proxy_async_smem();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> global
    ptx::mbar_wait(pv_done, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const int row = q0 + r;
    if (p.splits > 1 && row < p.Lq)
      p.part_ml[(static_cast<int64_t>(blockIdx.z) * p.Lq + row) * p.heads + head] = make_float2(m, l);
#pragma unroll 1
    for (int c = 0; c < kHD / 32; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + lane_sel + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < p.Lq) {
        if (p.splits > 1) {
          float* dst = p.part_o + (static_cast<int64_t>(blockIdx.z) * p.Lq + row) * (static_cast<int64_t>(p.heads) * kHD) + head * kHD + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l,
                                                              __uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        } else {
          __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + head * kHD + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + i) = w;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// Merge of the split-KV partials: out = sum_s w_s O_s / sum_s w_s with w_s = l_s * 2^(m_s - max_s m_s).
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, int splits,
                                                           int Lq, int heads, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int groups_per_row = heads * (kHD / 8);
  const int64_t total = static_cast<int64_t>(Lq) * groups_per_row;
  const int64_t W = static_cast<int64_t>(heads) * kHD;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / groups_per_row), g = static_cast<int>(i % groups_per_row);
    const int head = g / (kHD / 8);
    float mmax = -INFINITY;
    for (int s = 0; s < splits; ++s) mmax = fmaxf(mmax, part_ml[(static_cast<int64_t>(s) * Lq + row) * heads + head].x);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float2 ml = part_ml[(static_cast<int64_t>(s) * Lq + row) * heads + head];
      const float w = ml.y * ptx::ex2_approx(ml.x - mmax);
      wsum += w;
      float v[8];
      ptx::ld_nc_v8_f32(part_o + (static_cast<int64_t>(s) * Lq + row) * W + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, v[j], acc[j]);
    }
    const float inv = 1.0f / wsum;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo + g * 8) = pack_bf16x8(acc);
  }
}

// scratch for the split-KV partials: grown on demand (first use happens in an eager call, before any graph capture)
static float* g_part_o = nullptr;
static float2* g_part_ml = nullptr;
static size_t g_part_elems = 0;

static int32_t ensure_split_scratch(int splits, int Lq, int heads) {
  const size_t need = static_cast<size_t>(splits) * Lq * heads * kHD;
  if (need <= g_part_elems) return MC_OK;
  // the previous (smaller) buffers are deliberately not freed: captured CUDA graphs may still point at them
  g_part_o = nullptr;
  g_part_ml = nullptr;
  g_part_elems = 0;
  cudaError_t e = cudaMalloc(&g_part_o, need * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&g_part_ml, static_cast<size_t>(splits) * Lq * heads * sizeof(float2));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(split-KV scratch)");
  g_part_elems = need;
  return MC_OK;
}

}  // namespace mc

extern "C" int32_t mc_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* vt, int64_t ldvt, void* out,
                               int64_t ldo, int32_t Lq, int32_t Lk, int32_t heads, float scale, void* stream) {
  MC_CHECK_ARG(q && k && vt && out, "mc_attn_fwd: null pointer");
  MC_CHECK_ARG(Lq >= 1 && Lk >= 1 && heads >= 1, "mc_attn_fwd: Lq=%d Lk=%d heads=%d", Lq, Lk, heads);
  const int64_t width = static_cast<int64_t>(heads) * mc::kHD;
  MC_CHECK_ARG(ldq >= width && ldk >= width && ldo >= width && ldvt >= Lk, "mc_attn_fwd: leading dimensions too small");
  MC_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldvt % 8 == 0 && ldo % 8 == 0, "mc_attn_fwd: leading dimensions must be multiples of 8");
  MC_CHECK_ARG(mc::aligned16(q) && mc::aligned16(k) && mc::aligned16(vt) && mc::aligned16(out), "mc_attn_fwd: pointers must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  int32_t rc = mc::make_tmap_bf16_2d(&tq, q, static_cast<uint64_t>(Lq), static_cast<uint64_t>(width), static_cast<uint64_t>(ldq), mc::kBQ, 64);
  if (rc) return rc;
  rc = mc::make_tmap_bf16_2d(&tk, k, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldk), mc::kBKV, 64);
  if (rc) return rc;
  rc = mc::make_tmap_bf16_2d(&tv, vt, static_cast<uint64_t>(width), static_cast<uint64_t>(Lk), static_cast<uint64_t>(ldvt), mc::kHD, mc::kBKV);
  if (rc) return rc;
  static int variant = -1;  // MC_ATTN_VARIANT: 0 = P via smem, 1 = P in TMEM (default)
  if (variant < 0) {
    const char* ev = getenv("MC_ATTN_VARIANT");
    const int v = (ev && ev[0] == '0') ? 0 : 1;
    cudaError_t e = cudaFuncSetAttribute(mc::attn_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mc::kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mc::attn_fwd_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mc::kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mc::attn_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mc::kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mc::attn_fwd_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mc::kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mc::attn_fwd_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mc::kAttnSmem);
    if (e != cudaSuccess) return mc::cuda_fail(e, "cudaFuncSetAttribute(attn smem)");
    variant = v;
  }
  // Work split. One CTA = one 128-row query tile of one head; with fewer than two waves of CTAs (2 per SM) the last,
  // partially filled wave dominates (token-sharded runs: 4095 rows x 12 heads = 384 CTAs on 296 slots), so the KV range is
  // split across blockIdx.z until there are about three waves, and the partial softmaxes are merged by a second kernel.
  const int q_tiles = (Lq + mc::kBQ - 1) / mc::kBQ;
  const int total_tiles = (Lk + mc::kBKV - 1) / mc::kBKV;
  const double waves = static_cast<double>(q_tiles) * heads / (2.0 * mc::num_sms());
  // MC_ATTN_POLY=k (0..3): k/4 of the softmax exponentials on the FMA pipe instead of the MUFU. Read per call (tools/bench_poly.py).
  const char* ep = getenv("MC_ATTN_POLY");
  int poly = ep ? atoi(ep) : mc::kDefaultPoly;
  if (poly < 0 || poly > 3) poly = mc::kDefaultPoly;
  // MC_ATTN_SPLITS=n forces n splits (1 = never split); unset/0 = choose by wave fill. Read per call (tools/bench_split.py).
  const char* es = getenv("MC_ATTN_SPLITS");
  int forced_splits = es ? atoi(es) : 0;
  if (forced_splits < 0 || forced_splits > 16) forced_splits = 0;
  int splits = 1;
  if (forced_splits > 0) {
    splits = forced_splits;
  } else if (waves < 4.0 && total_tiles >= 16) {
    // smallest split count whose last wave is at least 92 % full, else the fullest
    double best = 0.0;
    for (int sp = 1; sp <= 8 && sp * 8 <= total_tiles; ++sp) {
      const double w = waves * sp, fill = w / static_cast<double>(static_cast<int64_t>(w + 0.999999));
      if (fill > best + 1e-9) best = fill, splits = sp;
      if (fill >= 0.92) break;
    }
  }
  if (splits > total_tiles) splits = total_tiles;
  if (splits > 1) {
    const int per = (total_tiles + splits - 1) / splits;
    splits = (total_tiles + per - 1) / per;  // no empty split
  }
  if (splits > 1) {
    MC_CHECK_ARG(ldo % 8 == 0, "mc_attn_fwd: ldo must be a multiple of 8");
    rc = mc::ensure_split_scratch(splits, Lq, heads);
    if (rc) return rc;
  }
  mc::AttnParams p{Lq, Lk, heads, splits, mc::g_part_o, mc::g_part_ml, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out), ldo};
  dim3 grid(q_tiles, heads, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant == 1)
    switch (poly) {
      case 1: mc::attn_fwd_kernel<1, 1><<<grid, mc::kAttnThreads, mc::kAttnSmem, st>>>(tq, tk, tv, p); break;
      case 2: mc::attn_fwd_kernel<1, 2><<<grid, mc::kAttnThreads, mc::kAttnSmem, st>>>(tq, tk, tv, p); break;
      case 3: mc::attn_fwd_kernel<1, 3><<<grid, mc::kAttnThreads, mc::kAttnSmem, st>>>(tq, tk, tv, p); break;
      default: mc::attn_fwd_kernel<1, 0><<<grid, mc::kAttnThreads, mc::kAttnSmem, st>>>(tq, tk, tv, p); break;
    }
  else
    mc::attn_fwd_kernel<0, 0><<<grid, mc::kAttnThreads, mc::kAttnSmem, st>>>(tq, tk, tv, p);
  MC_CHECK_LAUNCH("attn_fwd_kernel launch");
  if (splits > 1) {
    const int64_t total = static_cast<int64_t>(Lq) * heads * (mc::kHD / 8);
    const int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(mc::num_sms()) * 8;
    mc::attn_combine_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, st>>>(mc::g_part_o, mc::g_part_ml, splits, Lq, heads,
                                                                                      static_cast<__nv_bfloat16*>(out), ldo);
    MC_CHECK_LAUNCH("attn_combine_kernel launch");
  }
  return MC_OK;
}
