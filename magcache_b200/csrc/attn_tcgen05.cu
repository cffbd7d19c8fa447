// Non-causal flash-attention forward on tcgen05/TMEM, head_dim 128 — the self-/cross-attention of the Wan DiT block
// (reference call chain MagCache4Wan2.1/magcache_generate.py:297-298 -> WanSelfAttention.forward ->
// flash_attention / SDPA, upstream wan/modules/{model,attention}.py) and the joint attention of the MMDiT blocks
// (MagCache4FLUX/magcache_flux.py:343-425, MagCache4HunyuanVideo/magcache_sample_video.py:108-140).
//
// q [Lq, H*128], k [Lk, H*128], v [Lk, H*128] are all ROW-MAJOR bf16 views (row pitch allowed: they are column slices of the
// fused q|k|v projection buffer). K is the B operand of S = Q K^T in K-major form; V is the B operand of O += P V in MN-major
// form (d contiguous) — the same [rows][64 bf16] 128-byte-swizzled TMA boxes for both, no transposed copy of V anywhere.
//
// Two kernels, one contract:
//   attn_long_kernel  (Lk >= kLongMinLk): one CTA = TWO 128-row query tiles of one head (256 query rows), KV streamed in
//     128-row tiles, one CTA per SM (192 KB smem, all 512 TMEM columns: S0 S1 O0 O1). Two softmax warpgroups, one per query
//     tile; the MMA warp serves them alternately (PV0, S0', PV1, S1', ...), so while warpgroup 0 exponentiates tile j the
//     tensor pipe runs PV1(j-1) and S1(j) — 1024 tensor cycles per window. The exponentials alone would fill that window
//     (16 MUFU results per clock per SM against 4096 MACs per clock: one ex2 per 256 MACs), so a fixed fraction of them is
//     evaluated on the FMA pipe (ptx::ex2_emul_pair); the S MMA runs at N = 128 (128 B of smem operands per clock, the
//     UMMA fetch limit — at N = 64 it was 192 B/clk and fetch-bound) and every K / V tile is fetched once per 256 rows.
//   attn_short_kernel (short key sequences: text / image cross-attention, refiner): one CTA = one 128-row query tile, 64-row
//     KV tiles, two CTAs co-resident per SM so that one CTA's prologue / epilogue overlaps the other's main loop.
// Both: thread = query row (TMEM lane) in the softmax warps, online softmax in the exp2 domain with lazy rescaling of O
// (only when the running max grows by more than 2^8), P -> bf16 -> tcgen05.st over the consumed S columns, PV MMA with A
// from TMEM. The units of a partially filled last wave are split over the KV range and merged by attn_combine_kernel.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

#ifndef MC_ATTN_EMU_DEFAULT
#define MC_ATTN_EMU_DEFAULT 2  // eighths of the softmax exponentials on the FMA pipe (long kernel); MC_ATTN_EMU overrides per call
#endif

namespace mc {

constexpr int kHD = 128;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kLongMinLk = 1024;

struct AttnParams {
  int Lq, Lk, heads;
  // Work decomposition (plan_attention): a unit = one query block of one head. CTAs [0, full_units) each own a whole unit;
  // the remaining `tail_units` units (the part of the grid that would not fill a last wave) are split `tail_split` ways over
  // the KV range, CTA full_units + u * tail_split + s owning split s of tail unit u and writing a partial result.
  int q_blocks, full_units, tail_units, tail_split;
  float* part_o;   // [tail_split][tail_units][rows per CTA][128] fp32, each split's output normalised by its own row sum
  float2* part_ml; // [tail_split][tail_units][rows per CTA]      (row max in the scaled log2 domain, row sum)
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  int64_t ldo;
  // KV tile order. Token-sharded runs consume the LOCAL keys first and the peers' keys in arrival order: tile j of this CTA is
  // global tile (tile_rot + j) mod total. seg_flags != nullptr: before a tile that touches rows of source segment s
  // (rows [s*seg_rows, (s+1)*seg_rows)) is loaded, seg_flags[s] must have reached seg_epoch (written by the peer copy).
  int tile_rot;
  const uint32_t* seg_flags;
  const uint32_t* seg_epoch;  // device memory: the epoch the flags must have reached (read at run time, so graph replays work)
  int seg_rows;
};

struct WorkItem {
  int q_block, head;
  int split, nsplit;  // nsplit > 1: this CTA covers KV tiles [split * per, ...) of its unit and writes a partial
  int tail_unit;
};

__device__ __forceinline__ WorkItem work_item(const AttnParams& p) {
  WorkItem w;
  int c = blockIdx.x, unit;
  if (c < p.full_units) {
    unit = c, w.split = 0, w.nsplit = 1, w.tail_unit = 0;
  } else {
    c -= p.full_units;
    w.tail_unit = c / p.tail_split;
    w.split = c - w.tail_unit * p.tail_split;
    w.nsplit = p.tail_split;
    unit = p.full_units + w.tail_unit;
  }
  // query blocks of one head are adjacent in launch order: the CTAs resident together stream the same K/V through L2
  w.head = unit / p.q_blocks;
  w.q_block = unit - w.head * p.q_blocks;
  return w;
}

// P = 2^x for four consecutive elements (two packed pairs): MUFU for a pair unless its bit in EMU_MASK is set, in which case the
// pair goes through the FMA-pipe polynomial. `pair` is the running pair index (mod 8 selects the mask bit).
template <uint32_t EMU_MASK>
__device__ __forceinline__ void exp2_pair(float& a, float& b, int pair) {
  if ((EMU_MASK >> (pair & 7)) & 1u) {
    ptx::ex2_emul_pair(a, b);
  } else {
    a = ptx::ex2_approx(a);
    b = ptx::ex2_approx(b);
  }
}

// Ragged last KV tile: the score columns of keys past Lk are overwritten with -inf IN TMEM before the softmax reads the tile.
// Kept out of line and on the TMEM side on purpose: as a register-level `if (col >= valid) s = -inf` the compiler if-converts
// the test into a compare + select per element of EVERY tile (+40 % instructions in the softmax loop); this runs once per CTA.
__device__ __noinline__ void mask_padding_columns(uint32_t tmem_row, int valid, int width) {
  for (int h = 0; h < width; h += 32) {
    if (h + 32 <= valid) continue;
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(tmem_row + h, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (h + c >= valid) v[c] = 0xff800000u;  // -inf
    ptx::tmem_st_32x32b_x32(tmem_row + h, v);
  }
  ptx::tmem_st_wait();
}

// token-sharded runs: block until every source segment a KV tile touches has landed (flag written by the peer's copy stream
// AFTER the segment's data, same stream). Called by the single TMA-issuing thread; the acquire load orders the flag before the
// tile loads in the generic proxy, the proxy fence carries that order over to the async proxy the TMA reads through.
__device__ __forceinline__ void wait_segments(const AttnParams& p, int row0, int rows) {
  if (p.seg_flags == nullptr) return;
  const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(p.seg_epoch);
  const int last = min(row0 + rows, p.Lk) - 1;
  const int s0 = row0 / p.seg_rows, s1 = last / p.seg_rows;
  for (int s = s0; s <= s1; ++s) {
    const uint32_t* f = p.seg_flags + s;
    const long long t0 = clock64();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (static_cast<int32_t>(v - want) >= 0) break;
      __nanosleep(100);
      if (clock64() - t0 > MC_MBAR_TIMEOUT_CYCLES) {
        printf("attention: key segment %d never arrived (flag %u, epoch %u)\n", s, v, want);
        __trap();
      }
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

// =====================================================================================================================
// long-sequence kernel: 256 query rows per CTA, 128-row KV tiles
// =====================================================================================================================
namespace lk {
constexpr int kBQ = 128, kBKV = 128;
constexpr int kQTileBytes = kBQ * kHD * 2;  // 32 KB per query tile (two 64-column boxes)
constexpr int kKBytes = kBKV * kHD * 2;     // 32 KB (two boxes [128 kv x 64 hd])
constexpr int kVBytes = kBKV * kHD * 2;     // 32 KB (two boxes [128 kv x 64 d])
constexpr int kStages = 2;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + 2 * kQTileBytes;    // 64 KB
constexpr int kOffV = kOffK + kStages * kKBytes;  // +64 KB
constexpr int kOffBar = kOffV + kStages * kVBytes;  // 192 KB
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 384;  // warps 0-3 softmax of query tile 0, 4-7 of query tile 1, warp 8 TMA, warp 9 MMA, 10-11 idle
constexpr int kSoftmaxRegs = 208, kDataRegs = 80;  // setmaxnreg: the data-path warpgroup hands its registers to the softmax ones
constexpr int kTmemCols = 512;  // S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t (64 packed columns) overwrites S_t
}  // namespace lk

template <uint32_t EMU_MASK>
__global__ void __launch_bounds__(lk::kThreads, 1)
    attn_long_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using namespace lk;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  uint64_t* p_part = bars + 11;   // [2][4] per query tile and quarter, 128 arrivals: P of keys [32 q, 32 q + 32) of the tile is in TMEM
  uint64_t* pv_done = bars + 19;  // [2] per query tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WorkItem wi = work_item(p);
  const int q0 = wi.q_block * (2 * kBQ);
  const int head = wi.head;
  const int total_tiles = (p.Lk + kBKV - 1) / kBKV;
  const int per_split = (total_tiles + wi.nsplit - 1) / wi.nsplit;
  const int t0 = wi.split * per_split;                   // first KV tile (in rotated order) of this CTA
  const int n_tiles = min(per_split, total_tiles - t0);  // >= 1, guaranteed by the host
  auto tile_of = [&](int j) {  // global KV tile index of this CTA's j-th tile
    int t = p.tile_rot + t0 + j;
    return t >= total_tiles ? t - total_tiles : t;
  };

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_long_kernel: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_v);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      for (int q = 0; q < 4; ++q) ptx::mbar_init(&p_part[s * 4 + q], 128);
      ptx::mbar_init(&pv_done[s], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 9) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 10) {
    ptx::setmaxnreg_dec<kDataRegs>();  // idle warps of the data-path warpgroup
  } else if (warp == 8) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    // (whole warp in uniform control flow, the copies issued by one elected lane: see ptx::elect_one)
    ptx::setmaxnreg_dec<kDataRegs>();
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, 2 * kQTileBytes);
      for (int t = 0; t < 2; ++t) {
        ptx::tma_load_2d(smem + kOffQ + t * kQTileBytes, &tmap_q, q_full, head * kHD, q0 + t * kBQ);
        ptx::tma_load_2d(smem + kOffQ + t * kQTileBytes + kQTileBytes / 2, &tmap_q, q_full, head * kHD + 64, q0 + t * kBQ);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv0 = tile_of(j) * kBKV;
      if (p.seg_flags != nullptr) {
        if (lane == 0) wait_segments(p, kv0, kBKV);
        __syncwarp();
      }
      ptx::mbar_wait(&k_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&k_full[s], kKBytes);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes, &tmap_k, &k_full[s], head * kHD, kv0);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes + kKBytes / 2, &tmap_k, &k_full[s], head * kHD + 64, kv0);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&v_full[s], kVBytes);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes, &tmap_v, &v_full[s], head * kHD, kv0);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes + kVBytes / 2, &tmap_v, &v_full[s], head * kHD + 64, kv0);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    // All 32 lanes walk the loop and wait on the barriers; the MMAs and commits are issued by the elected lane.
    ptx::setmaxnreg_dec<kDataRegs>();
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(kBQ, kBKV);      // S: 128 x 128, both operands K-major
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32_bmn(kBQ, kHD);   // O: 128 x 128, B = V tile MN-major
    const uint32_t q_base = ptx::smem_u32(smem + kOffQ), k_base = ptx::smem_u32(smem + kOffK), v_base = ptx::smem_u32(smem + kOffV);
    auto issue_s = [&](int t, int j) {  // S_t(j) = Q_t K_j^T   (elected lane only)
      const uint64_t da0 = ptx::umma_desc_sw128_kmajor(q_base + t * kQTileBytes);
      const uint64_t db0 = ptx::umma_desc_sw128_kmajor(k_base + (j & 1) * kKBytes);
#pragma unroll
      for (int kk = 0; kk < kHD / 16; ++kk) {
        // 64-column half (kk >> 2) is a separate TMA box kQTileBytes / 2 further; inside a box one K-step is 32 B (+2 in the >>4 field)
        const uint64_t off = static_cast<uint64_t>((kk >> 2) * (kQTileBytes / 2 / 16) + 2 * (kk & 3));
        ptx::umma_ss(tmem_base + t * 128, da0 + off, db0 + off, idesc_s, kk != 0 ? 1u : 0u);
      }
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&k_full[0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      issue_s(0, 0);
      ptx::umma_commit(&s_full[0]);
      issue_s(1, 0);
      ptx::umma_commit(&k_empty[0]);
      ptx::umma_commit(&s_full[1]);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const uint64_t dv0 = ptx::umma_desc_sw128_mnmajor(v_base + (j & 1) * kVBytes, kBKV * 128);  // LBO = one [128 kv x 64 d] box
      for (int t = 0; t < 2; ++t) {
        // O_t += P_t(j) V_j, A from TMEM (8 packed columns per K16 step), in four quarters: every 32 keys' probabilities are
        // signalled as soon as they are written, so most of the MMA runs while the softmax warps exponentiate the rest and only
        // the last quarter (plus the next S MMA) is left on the softmax -> PV -> S -> softmax chain of this query tile
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          ptx::mbar_wait(&p_part[t * 4 + part], j & 1);
          if (t == 0 && part == 0) ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = part * 2; kk < part * 2 + 2; ++kk)  // one K-step = 16 kv rows = 2048 B of the MN-major tile
              ptx::umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8, dv0 + static_cast<uint64_t>(kk * (2048 / 16)), idesc_o,
                           (j | kk) != 0 ? 1u : 0u);
            if (part == 3) {
              if (t == 1) ptx::umma_commit(&v_empty[j & 1]);
              ptx::umma_commit(&pv_done[t]);
            }
          }
          __syncwarp();
        }
        if (j + 1 < n_tiles) {
          if (t == 0) {
            ptx::mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
            ptx::tc_fence_after();
          }
          if (ptx::elect_one()) {
            issue_s(t, j + 1);  // overwrites S_t / P_t(j): ordered behind PV_t(j) by the tensor pipe
            if (t == 1) ptx::umma_commit(&k_empty[(j + 1) & 1]);
            ptx::umma_commit(&s_full[t]);  // fires when S_t(j+1) AND everything before it (PV_t(j)) has completed
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax warpgroups ------------------------------------------
    ptx::setmaxnreg_inc<kSoftmaxRegs>();
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * 128 + lane_sel;
    const uint32_t tmem_o = tmem_base + 256 + t * 128 + lane_sel;
    const uint64_t scale2 = ptx::pack_f32x2(p.scale_log2, p.scale_log2);
    float m = -INFINITY, l = 0.f;
    // the globally last KV tile is the only one that can hold padding columns (keys >= Lk): its position in this CTA's order
    int j_ragged = -1;
    if (p.Lk % kBKV != 0) {
      int jr = total_tiles - 1 - p.tile_rot - t0;
      if (jr < 0) jr += total_tiles;
      if (jr < n_tiles) j_ragged = jr;
    }

    for (int j = 0; j < n_tiles; ++j) {
      // s_full(j) also certifies that PV_t(j-1) has completed (commit semantics): O_t is quiescent until p_part[0](j) is signalled
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      if (j == j_ragged) mask_padding_columns(tmem_s, p.Lk - (total_tiles - 1) * kBKV, kBKV);  // warp-uniform, at most once per CTA
      uint32_t sreg[4][32];
#pragma unroll
      for (int h = 0; h < 4; ++h) ptx::tmem_ld_32x32b_x32(tmem_s + h * 32, sreg[h]);
      ptx::tmem_ld_wait();
      // 3-input max, 0.5 instruction per element, in EIGHT independent chains (two per 32-column chunk): the exponentials cannot
      // start before the row max is known, so the chain depth (8 dependent ops) is what this costs, not the instruction count
      float mxa[4], mxb[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        mxa[h] = -INFINITY, mxb[h] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          mxa[h] = ptx::max3(mxa[h], __uint_as_float(sreg[h][c]), __uint_as_float(sreg[h][c + 1]));
          mxb[h] = ptx::max3(mxb[h], __uint_as_float(sreg[h][c + 2]), __uint_as_float(sreg[h][c + 3]));
        }
      }
      const float mx = fmaxf(ptx::max3(mxa[0], mxa[1], mxa[2]), fmaxf(mxa[3], fmaxf(ptx::max3(mxb[0], mxb[1], mxb[2]), mxb[3])));
      const float m_new = fmaxf(m, mx * p.scale_log2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = m_new > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float factor = need ? ptx::ex2_approx(m - m_new) : 1.0f;
          if (need) {
            l *= factor;
            m = m_new;
          }
#pragma unroll 1
          for (int c = 0; c < kHD / 32; ++c) {
            uint32_t o[32];
            ptx::tmem_ld_32x32b_x32(tmem_o + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            ptx::tmem_st_32x32b_x32(tmem_o + c * 32, o);
          }
        }
      }
      // P = 2^(s*scale - m), row sum in fp32 before the bf16 rounding (as flash-attention does). Packed fp32x2 FMA / ADD;
      // 32 columns -> 16 packed words, stored over the consumed S columns right away.
      const uint64_t negm2 = ptx::pack_f32x2(-m, -m);
      uint64_t sum2a = 0ull, sum2b = 0ull;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float a0, a1, b0, b1;
          ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[h][c]), __uint_as_float(sreg[h][c + 1])), scale2, negm2), a0, a1);
          ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[h][c + 2]), __uint_as_float(sreg[h][c + 3])), scale2, negm2), b0, b1);
          exp2_pair<EMU_MASK>(a0, a1, (h * 32 + c) >> 1);
          exp2_pair<EMU_MASK>(b0, b1, ((h * 32 + c) >> 1) + 1);
          sum2a = ptx::add_f32x2(sum2a, ptx::pack_f32x2(a0, a1));
          sum2b = ptx::add_f32x2(sum2b, ptx::pack_f32x2(b0, b1));
          pk[c >> 1] = pack_bf16x2(a0, a1);
          pk[(c >> 1) + 1] = pack_bf16x2(b0, b1);
        }
        ptx::tmem_st_32x32b_x16(tmem_s + h * 16, pk);
        ptx::tmem_st_wait();  // keys [32 h, 32 h + 32) of this tile are in place: let that quarter of PV_t(j) go
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_part[t * 4 + h]);
      }
      float s0, s1, s2, s3;
      ptx::unpack_f32x2(sum2a, s0, s1);
      ptx::unpack_f32x2(sum2b, s2, s3);
      l += (s0 + s1) + (s2 + s3);
    }

    // ---- epilogue: O_t / l -> bf16 -> global (or the normalised fp32 partial of this split)
    ptx::mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const int row = q0 + t * kBQ + r;
    const bool partial = wi.nsplit > 1;
    const int64_t prow = (static_cast<int64_t>(wi.split) * p.tail_units + wi.tail_unit) * (2 * kBQ) + t * kBQ + r;
    if (partial) p.part_ml[prow] = make_float2(m, l);
#pragma unroll 1
    for (int c = 0; c < kHD / 32; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < p.Lq) {
        if (partial) {
          float* dst = p.part_o + prow * kHD + c * 32;
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
  if (warp == 9) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// short-sequence kernel: 128 query rows per CTA, 64-row KV tiles, two CTAs per SM
// =====================================================================================================================
namespace sk {
constexpr int kBQ = 128, kBKV = 64;
constexpr int kQBytes = kBQ * kHD * 2;   // 32 KB (two 64-column boxes of 16 KB)
constexpr int kKBytes = kBKV * kHD * 2;  // 16 KB (two boxes of 8 KB)
constexpr int kVBytes = kBKV * kHD * 2;  // 16 KB (two boxes [64 kv x 64 d] of 8 KB)
constexpr int kKVStages = 2;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kKVStages * kKBytes;
constexpr int kOffBar = kOffV + kKVStages * kVBytes;  // 96 KB
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 192;   // warps 0-3 softmax, warp 4 TMA, warp 5 MMA
constexpr int kTmemCols = 256;  // S0 [0,64) S1 [64,128) O [128,256)
}  // namespace sk

__global__ void __launch_bounds__(sk::kThreads, 2)
    attn_short_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using namespace sk;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WorkItem wi = work_item(p);
  const int q0 = wi.q_block * kBQ;
  const int head = wi.head;
  const int total_tiles = (p.Lk + kBKV - 1) / kBKV;
  const int per_split = (total_tiles + wi.nsplit - 1) / wi.nsplit;
  const int t0 = wi.split * per_split;                   // first (global) KV tile of this CTA
  const int n_tiles = min(per_split, total_tiles - t0);  // local tile count (>= 1, guaranteed by the host)

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_short_kernel: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_v);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
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
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, kQBytes);
      ptx::tma_load_2d(smem + kOffQ, &tmap_q, q_full, head * kHD, q0);
      ptx::tma_load_2d(smem + kOffQ + kQBytes / 2, &tmap_q, q_full, head * kHD + 64, q0);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv0 = (t0 + j) * kBKV;
      ptx::mbar_wait(&k_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&k_full[s], kKBytes);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes, &tmap_k, &k_full[s], head * kHD, kv0);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes + kKBytes / 2, &tmap_k, &k_full[s], head * kHD + 64, kv0);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&v_full[s], kVBytes);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes, &tmap_v, &v_full[s], head * kHD, kv0);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes + kVBytes / 2, &tmap_v, &v_full[s], head * kHD + 64, kv0);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer (elected lane, uniform control flow) --------------
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(kBQ, kBKV);     // 128 x 64
    constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32_bmn(kBQ, kHD);  // 128 x 128, B = V tile MN-major
    const uint32_t q_base = ptx::smem_u32(smem + kOffQ), k_base = ptx::smem_u32(smem + kOffK), v_base = ptx::smem_u32(smem + kOffV);
    auto issue_s = [&](int j) {  // elected lane only
      const int s = j & 1;
      const uint64_t da0 = ptx::umma_desc_sw128_kmajor(q_base), db0 = ptx::umma_desc_sw128_kmajor(k_base + s * kKBytes);
#pragma unroll
      for (int kk = 0; kk < kHD / 16; ++kk)
        ptx::umma_ss(tmem_base + s * kBKV, da0 + static_cast<uint64_t>((kk >> 2) * (kQBytes / 2 / 16) + 2 * (kk & 3)),
                     db0 + static_cast<uint64_t>((kk >> 2) * (kKBytes / 2 / 16) + 2 * (kk & 3)), idesc_s, kk != 0 ? 1u : 0u);
      ptx::umma_commit(&k_empty[s]);
      ptx::umma_commit(&s_full[s]);
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&k_full[0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) {
        const int t = j + 1;
        ptx::mbar_wait(&k_full[t & 1], (t >> 1) & 1);
        // S buffer t&1 was last used by tile t-2: its softmax drained it before p_full(t-2), which was waited before PV(t-2)
        // was issued, and PV(t-2) (the reader of P in that buffer) executes before this MMA on the in-order tensor pipe.
        ptx::tc_fence_after();
        if (ptx::elect_one()) issue_s(t);
        __syncwarp();
      }
      ptx::mbar_wait(p_full, j & 1);
      ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dv0 = ptx::umma_desc_sw128_mnmajor(v_base + (j & 1) * kVBytes, kBKV * 128);  // LBO = one [64 kv x 64 d] box
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk)
          ptx::umma_ts(tmem_o, tmem_base + (j & 1) * kBKV + kk * 8, dv0 + static_cast<uint64_t>(kk * (2048 / 16)), idesc_o, (j | kk) != 0 ? 1u : 0u);
        ptx::umma_commit(&v_empty[j & 1]);
        ptx::umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ softmax warps -----------------------------------------------
    const int r = warp * 32 + lane;  // row inside the Q tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      ptx::mbar_wait(&s_full[b], (j >> 1) & 1);
      ptx::tc_fence_after();
      const int valid = p.Lk - (t0 + j) * kBKV;  // columns >= valid are padding (only possible on the last tile)
      if (valid < kBKV) mask_padding_columns(tmem_base + lane_sel + b * kBKV, valid, kBKV);  // warp-uniform, at most once per CTA
      uint32_t sreg[2][32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_sel + b * kBKV, sreg[0]);
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_sel + b * kBKV + 32, sreg[1]);
      ptx::tmem_ld_wait();
      float mx0 = -INFINITY, mx1 = -INFINITY;
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
          a0 = ptx::ex2_approx(a0);
          a1 = ptx::ex2_approx(a1);
          b0 = ptx::ex2_approx(b0);
          b1 = ptx::ex2_approx(b1);
          sum2a = ptx::add_f32x2(sum2a, ptx::pack_f32x2(a0, a1));
          sum2b = ptx::add_f32x2(sum2b, ptx::pack_f32x2(b0, b1));
          packed[h * 16 + (c >> 1)] = pack_bf16x2(a0, a1);
          packed[h * 16 + (c >> 1) + 1] = pack_bf16x2(b0, b1);
        }
      float s0, s1, s2, s3;
      ptx::unpack_f32x2(sum2a, s0, s1);
      ptx::unpack_f32x2(sum2b, s2, s3);
      l += (s0 + s1) + (s2 + s3);
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
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> global
    ptx::mbar_wait(pv_done, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const int row = q0 + r;
    const bool partial = wi.nsplit > 1;
    const int64_t prow = (static_cast<int64_t>(wi.split) * p.tail_units + wi.tail_unit) * kBQ + r;
    if (partial) p.part_ml[prow] = make_float2(m, l);
#pragma unroll 1
    for (int c = 0; c < kHD / 32; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + lane_sel + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < p.Lq) {
        if (partial) {
          float* dst = p.part_o + prow * kHD + c * 32;
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

// Merge of the split-KV partials of the tail units: out = sum_s w_s O_s / sum_s w_s with w_s = l_s * 2^(m_s - max_s m_s).
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, int nsplit,
                                                           int tail_units, int full_units, int q_blocks, int rows_per_cta, int Lq,
                                                           __nv_bfloat16* __restrict__ out, int64_t ldo) {
  constexpr int kGroups = kHD / 8;
  const int64_t total = static_cast<int64_t>(tail_units) * rows_per_cta * kGroups;
  const int64_t split_stride = static_cast<int64_t>(tail_units) * rows_per_cta;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t prow = i / kGroups;
    const int g = static_cast<int>(i - prow * kGroups);
    const int tu = static_cast<int>(prow / rows_per_cta), r = static_cast<int>(prow - static_cast<int64_t>(tu) * rows_per_cta);
    const int unit = full_units + tu;
    const int head = unit / q_blocks, row = (unit - head * q_blocks) * rows_per_cta + r;
    if (row >= Lq) continue;
    float mmax = -INFINITY;
    for (int s = 0; s < nsplit; ++s) mmax = fmaxf(mmax, part_ml[s * split_stride + prow].x);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float2 ml = part_ml[s * split_stride + prow];
      const float w = ml.y * ptx::ex2_approx(ml.x - mmax);
      wsum += w;
      float v[8];
      ptx::ld_nc_v8_f32(part_o + (s * split_stride + prow) * kHD + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, v[j], acc[j]);
    }
    const float inv = 1.0f / wsum;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo + head * kHD + g * 8) = pack_bf16x8(acc);
  }
}

// ---- host side: work decomposition ----------------------------------------------------------------------------------
struct AttnPlan {
  bool long_kernel;
  int q_blocks, kv_tile, total_tiles, rows_per_cta;
  int full_units, tail_units, tail_split;  // see AttnParams
  size_t ws_bytes;                         // workspace for the split partials (0 when nothing is split)
  int grid() const { return full_units + tail_units * tail_split; }
};

static int env_int(const char* name, int lo, int hi, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return (v < lo || v > hi) ? dflt : v;
}

// One CTA = 256 (long kernel, 1 CTA/SM) or 128 (short kernel, 2 CTAs/SM) query rows of one head, every unit the same length, so
// the grid runs in rounds of `slots` CTAs and a partially filled last round costs a whole one (32760 rows x 12 heads = 1536
// units on 148 SMs: 10.4 rounds of work take 11; a token-sharded rank's 4095 rows: 192 units, 1.3 rounds take 2). The units
// of that last round are therefore split over the KV range, floor(slots / units) ways, so that they fill the SMs once at a
// fraction of the length; a second kernel merges their partial softmaxes. Units of the full rounds are never split.
static AttnPlan plan_attention(int Lq, int Lk, int heads, bool need_long = false) {
  AttnPlan pl;
  const int forced_splits = env_int("MC_ATTN_SPLITS", 1, 16, 0);  // MC_ATTN_SPLITS=n: every unit split n ways (1 = never split)
  const int kernel_sel = env_int("MC_ATTN_KERNEL", 0, 2, 0);  // 0 = by Lk, 1 = short, 2 = long (tests / A-B timing)
  pl.long_kernel = need_long || kernel_sel == 2 || (kernel_sel == 0 && Lk >= kLongMinLk);  // rotated / flag-gated key order: long kernel only
  pl.rows_per_cta = pl.long_kernel ? 2 * lk::kBQ : sk::kBQ;
  pl.kv_tile = pl.long_kernel ? lk::kBKV : sk::kBKV;
  pl.q_blocks = (Lq + pl.rows_per_cta - 1) / pl.rows_per_cta;
  pl.total_tiles = (Lk + pl.kv_tile - 1) / pl.kv_tile;
  const int64_t units = static_cast<int64_t>(pl.q_blocks) * heads;
  const int slots = (pl.long_kernel ? 1 : 2) * num_sms();
  const int min_tiles_per_split = pl.long_kernel ? 4 : 8;
  int full = static_cast<int>(units), tail = 0, split = 1;
  if (forced_splits > 0) {
    if (forced_splits > 1) full = 0, tail = static_cast<int>(units), split = forced_splits;
  } else {
    const int rem = static_cast<int>(units % slots);
    const int by_tiles = pl.total_tiles / min_tiles_per_split;
    int f = rem > 0 ? slots / rem : 1;
    if (f > 8) f = 8;
    if (f > by_tiles) f = by_tiles;
    if (f >= 2) full = static_cast<int>(units) - rem, tail = rem, split = f;
  }
  if (split > pl.total_tiles) split = pl.total_tiles;
  if (split > 1) {
    const int per = (pl.total_tiles + split - 1) / split;
    split = (pl.total_tiles + per - 1) / per;  // no empty split
  }
  if (split <= 1) full = static_cast<int>(units), tail = 0, split = 1;
  pl.full_units = full, pl.tail_units = tail, pl.tail_split = split;
  pl.ws_bytes = tail > 0 ? static_cast<size_t>(split) * tail * pl.rows_per_cta * (kHD * sizeof(float) + sizeof(float2)) : 0;
  return pl;
}

}  // namespace mc

extern "C" int32_t mc_attn_workspace_bytes(int32_t Lq, int32_t Lk, int32_t heads, int64_t* bytes_out) {
  MC_CHECK_ARG(bytes_out != nullptr && Lq >= 1 && Lk >= 1 && heads >= 1, "mc_attn_workspace_bytes: bad arguments");
  // the larger of the two decompositions the launcher may pick (the rotated / flag-gated form always takes the long kernel)
  const size_t a = mc::plan_attention(Lq, Lk, heads, false).ws_bytes, b = mc::plan_attention(Lq, Lk, heads, true).ws_bytes;
  *bytes_out = static_cast<int64_t>(a > b ? a : b);
  return MC_OK;
}

extern "C" int32_t mc_attn_fwd_ex(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                  int64_t ldo, int32_t Lq, int32_t Lk, int32_t heads, float scale, void* workspace,
                                  int64_t workspace_bytes, int32_t first_key_row, const uint32_t* seg_flags, const uint32_t* seg_epoch,
                                  int32_t seg_rows, void* stream) {
  MC_CHECK_ARG(q && k && v && out, "mc_attn_fwd: null pointer");
  MC_CHECK_ARG(Lq >= 1 && Lk >= 1 && heads >= 1, "mc_attn_fwd: Lq=%d Lk=%d heads=%d", Lq, Lk, heads);
  const int64_t width = static_cast<int64_t>(heads) * mc::kHD;
  MC_CHECK_ARG(ldq >= width && ldk >= width && ldo >= width && ldv >= width, "mc_attn_fwd: leading dimensions too small");
  MC_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "mc_attn_fwd: leading dimensions must be multiples of 8");
  MC_CHECK_ARG(mc::aligned16(q) && mc::aligned16(k) && mc::aligned16(v) && mc::aligned16(out), "mc_attn_fwd: pointers must be 16-byte aligned");
  MC_CHECK_ARG(first_key_row >= 0 && first_key_row < Lk, "mc_attn_fwd: first_key_row=%d outside [0, %d)", first_key_row, Lk);
  MC_CHECK_ARG(seg_flags == nullptr || (seg_rows >= 1 && seg_epoch != nullptr), "mc_attn_fwd: seg_rows=%d / null epoch", seg_rows);
  const mc::AttnPlan pl = mc::plan_attention(Lq, Lk, heads, first_key_row != 0 || seg_flags != nullptr);
  MC_CHECK_ARG(static_cast<int64_t>(pl.q_blocks) * heads * 8 < INT32_MAX, "mc_attn_fwd: grid too large");
  if (pl.tail_units > 0) {
    MC_CHECK_ARG(workspace != nullptr && workspace_bytes >= static_cast<int64_t>(pl.ws_bytes) && (reinterpret_cast<uintptr_t>(workspace) & 31u) == 0,
                 "mc_attn_fwd: split-KV needs a 32-byte aligned workspace of %lld bytes (mc_attn_workspace_bytes), got %lld",
                 static_cast<long long>(pl.ws_bytes), static_cast<long long>(workspace_bytes));
  }
  const int q_box = pl.long_kernel ? mc::lk::kBQ : mc::sk::kBQ;
  CUtensorMap tq, tk, tv;
  int32_t rc = mc::make_tmap_bf16_2d(&tq, q, static_cast<uint64_t>(Lq), static_cast<uint64_t>(width), static_cast<uint64_t>(ldq), q_box, 64);
  if (rc) return rc;
  rc = mc::make_tmap_bf16_2d(&tk, k, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldk), pl.kv_tile, 64);
  if (rc) return rc;
  rc = mc::make_tmap_bf16_2d(&tv, v, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldv), pl.kv_tile, 64);
  if (rc) return rc;

  float* part_o = static_cast<float*>(workspace);
  const size_t part_rows = static_cast<size_t>(pl.tail_split) * pl.tail_units * pl.rows_per_cta;
  float2* part_ml = pl.tail_units > 0 ? reinterpret_cast<float2*>(part_o + part_rows * mc::kHD) : nullptr;
  mc::AttnParams p{};
  p.Lq = Lq, p.Lk = Lk, p.heads = heads;
  p.q_blocks = pl.q_blocks, p.full_units = pl.full_units, p.tail_units = pl.tail_units, p.tail_split = pl.tail_split;
  p.part_o = pl.tail_units > 0 ? part_o : nullptr, p.part_ml = part_ml;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out), p.ldo = ldo;
  // start with the first tile that lies entirely inside the caller's own (already resident) rows; the tile straddling the segment
  // boundary before it comes last in the rotated order
  p.tile_rot = ((first_key_row + pl.kv_tile - 1) / pl.kv_tile) % pl.total_tiles;
  p.seg_flags = seg_flags, p.seg_epoch = seg_epoch, p.seg_rows = seg_rows > 0 ? seg_rows : Lk;

  dim3 grid(pl.grid(), 1, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl.long_kernel) {
    static mc::PerDeviceOnce once[4];
    // MC_ATTN_EMU = eighths of the exponentials evaluated on the FMA pipe: 0, 2 (25 %), 3 (37.5 %), 4 (50 %)
    const int emu = mc::env_int("MC_ATTN_EMU", 0, 4, MC_ATTN_EMU_DEFAULT);
#define MC_LAUNCH_LONG(MASK, IDX)                                                                                             \
  do {                                                                                                                        \
    rc = mc::set_max_smem_once(mc::attn_long_kernel<MASK>, mc::lk::kSmem, once[IDX], "cudaFuncSetAttribute(attn long smem)"); \
    if (rc) return rc;                                                                                                        \
    mc::attn_long_kernel<MASK><<<grid, mc::lk::kThreads, mc::lk::kSmem, st>>>(tq, tk, tv, p);                                 \
  } while (0)
    switch (emu) {
      case 2: MC_LAUNCH_LONG(0x88u, 1); break;
      case 3: MC_LAUNCH_LONG(0x92u, 2); break;
      case 4: MC_LAUNCH_LONG(0xAAu, 3); break;
      default: MC_LAUNCH_LONG(0x00u, 0); break;
    }
#undef MC_LAUNCH_LONG
    MC_CHECK_LAUNCH("attn_long_kernel launch");
  } else {
    static mc::PerDeviceOnce once;
    rc = mc::set_max_smem_once(mc::attn_short_kernel, mc::sk::kSmem, once, "cudaFuncSetAttribute(attn short smem)");
    if (rc) return rc;
    mc::attn_short_kernel<<<grid, mc::sk::kThreads, mc::sk::kSmem, st>>>(tq, tk, tv, p);
    MC_CHECK_LAUNCH("attn_short_kernel launch");
  }
  if (pl.tail_units > 0) {
    const int64_t total = static_cast<int64_t>(pl.tail_units) * pl.rows_per_cta * (mc::kHD / 8);
    const int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(mc::num_sms()) * 8;
    mc::attn_combine_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, st>>>(
        p.part_o, p.part_ml, pl.tail_split, pl.tail_units, pl.full_units, pl.q_blocks, pl.rows_per_cta, Lq, static_cast<__nv_bfloat16*>(out), ldo);
    MC_CHECK_LAUNCH("attn_combine_kernel launch");
  }
  return MC_OK;
}

extern "C" int32_t mc_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                               int64_t ldo, int32_t Lq, int32_t Lk, int32_t heads, float scale, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  return mc_attn_fwd_ex(q, ldq, k, ldk, v, ldv, out, ldo, Lq, Lk, heads, scale, workspace, workspace_bytes, 0, nullptr, nullptr, 0, stream);
}
