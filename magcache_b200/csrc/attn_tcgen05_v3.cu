// Flash-attention forward, 256-row CTAs: two 128-row query tiles per CTA share every K / V^T tile, KV streamed in 128-row
// tiles, one CTA per SM (192 KB smem, all 512 TMEM columns). Same math and operand layouts as attn_tcgen05.cu (see there
// for the reference call chain); what changes is the work decomposition:
//   * 128-wide KV tiles: half as many softmax round trips (mbarrier wake-up, tcgen05.ld, tcgen05.st + wait, fence, arrive)
//     per key, and the S MMA runs at N = 128 (one smem operand byte per 2x the MACs of N = 64)
//   * each K / V^T tile is fetched once for 256 query rows (half the TMA / L2 traffic per FLOP)
//   * the two softmax warpgroups are served alternately by the MMA warp (PV0, S0', PV1, S1', ...), so one warpgroup's
//     wait-for-S gap is covered by the other's exponentials
// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t (bf16, 64 packed columns) overwrites the consumed S_t.
//   warps 0-3 softmax of query tile 0, warps 4-7 softmax of query tile 1 (thread = query row), warp 8 TMA, warp 9 MMA
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {
namespace v3 {

constexpr int kBQ = 128, kBKV = 128, kHD = 128;
constexpr int kQTileBytes = kBQ * kHD * 2;   // 32 KB per query tile (two 64-column boxes)
constexpr int kKBytes = kBKV * kHD * 2;      // 32 KB (two boxes [128 kv x 64 hd])
constexpr int kVBytes = kHD * kBKV * 2;      // 32 KB (two boxes [128 d x 64 kv])
constexpr int kStages = 2;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + 2 * kQTileBytes;           // 64 KB
constexpr int kOffV = kOffK + kStages * kKBytes;         // +64 KB
constexpr int kOffBar = kOffV + kStages * kVBytes;       // 192 KB
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct Params {
  int Lq, Lk, heads;
  float scale_log2;
  __nv_bfloat16* out;
  int64_t ldo;
};

__global__ void __launch_bounds__(kThreads, 1)
    attn_fwd_kernel_v3(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_vt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  uint64_t* p_full = bars + 11;   // [2] per query tile, 128 arrivals
  uint64_t* pv_done = bars + 13;  // [2] per query tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kBQ);
  const int head = blockIdx.y;
  const int n_tiles = (p.Lk + kBKV - 1) / kBKV;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_fwd_kernel_v3: dynamic smem base not 1024-aligned\n");
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
      ptx::mbar_init(&p_full[s], 128);
      ptx::mbar_init(&pv_done[s], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 9) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, 2 * kQTileBytes);
      for (int t = 0; t < 2; ++t) {
        ptx::tma_load_2d(smem + kOffQ + t * kQTileBytes, &tmap_q, q_full, head * kHD, q0 + t * kBQ);
        ptx::tma_load_2d(smem + kOffQ + t * kQTileBytes + kQTileBytes / 2, &tmap_q, q_full, head * kHD + 64, q0 + t * kBQ);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[s], kKBytes);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes, &tmap_k, &k_full[s], head * kHD, j * kBKV);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes + kKBytes / 2, &tmap_k, &k_full[s], head * kHD + 64, j * kBKV);
        ptx::mbar_wait(&v_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], kVBytes);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes, &tmap_vt, &v_full[s], j * kBKV, head * kHD);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes + kVBytes / 2, &tmap_vt, &v_full[s], j * kBKV + 64, head * kHD);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kBQ, 128);  // both MMAs are 128 x 128 (x K16)
      auto issue_s = [&](int t, int j) {  // S_t(j) = Q_t K_j^T
        const uint32_t q_addr = ptx::smem_u32(smem + kOffQ + t * kQTileBytes);
        const uint32_t k_addr = ptx::smem_u32(smem + kOffK + (j & 1) * kKBytes);
#pragma unroll
        for (int kk = 0; kk < kHD / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_sw128_kmajor(q_addr + (kk >> 2) * (kQTileBytes / 2)) + 2 * (kk & 3);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(k_addr + (kk >> 2) * (kKBytes / 2)) + 2 * (kk & 3);
          ptx::umma_ss(tmem_base + t * 128, da, db, idesc, kk != 0 ? 1u : 0u);
        }
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0, 0);
      ptx::umma_commit(&s_full[0]);
      issue_s(1, 0);
      ptx::umma_commit(&k_empty[0]);
      ptx::umma_commit(&s_full[1]);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t v_addr = ptx::smem_u32(smem + kOffV + (j & 1) * kVBytes);
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&p_full[t], j & 1);
          if (t == 0) ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kBKV / 16; ++kk) {  // O_t += P_t(j) V_j : A from TMEM (8 packed columns per K16 step)
            const uint64_t db = ptx::umma_desc_sw128_kmajor(v_addr + (kk >> 2) * (kVBytes / 2)) + 2 * (kk & 3);
            ptx::umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8, db, idesc, (j | kk) != 0 ? 1u : 0u);
          }
          if (t == 1) ptx::umma_commit(&v_empty[j & 1]);
          ptx::umma_commit(&pv_done[t]);
          if (j + 1 < n_tiles) {
            if (t == 0) {
              ptx::mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
              ptx::tc_fence_after();
            }
            issue_s(t, j + 1);  // overwrites S_t / P_t(j): ordered behind PV_t(j) by the tensor pipe
            if (t == 1) ptx::umma_commit(&k_empty[(j + 1) & 1]);
            ptx::umma_commit(&s_full[t]);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax warpgroups ------------------------------------------
    const int t = warp >> 2;                       // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * 128 + lane_sel;
    const uint32_t tmem_o = tmem_base + 256 + t * 128 + lane_sel;
    const uint64_t scale2 = ptx::pack_f32x2(p.scale_log2, p.scale_log2);
    float m = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      uint32_t sreg[4][32];
#pragma unroll
      for (int h = 0; h < 4; ++h) ptx::tmem_ld_32x32b_x32(tmem_s + h * 32, sreg[h]);
      ptx::tmem_ld_wait();
      const int valid = p.Lk - j * kBKV;
      if (valid < kBKV) {  // warp-uniform, at most once per CTA
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (h * 32 + c >= valid) sreg[h][c] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int h = 0; h < 4; ++h)
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          mx0 = ptx::max3(mx0, __uint_as_float(sreg[h][c]), __uint_as_float(sreg[h][c + 1]));
          mx1 = ptx::max3(mx1, __uint_as_float(sreg[h][c + 2]), __uint_as_float(sreg[h][c + 3]));
        }
      const float m_new = fmaxf(m, fmaxf(mx0, mx1) * p.scale_log2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = m_new > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          ptx::mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t quiescent
          ptx::tc_fence_after();
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
          ptx::tmem_st_wait();
        }
      }
      const uint64_t negm2 = ptx::pack_f32x2(-m, -m);
      uint64_t sum2a = 0ull, sum2b = 0ull;
#pragma unroll
      for (int h = 0; h < 4; ++h) {  // 32 columns -> 16 packed words, stored over the consumed S columns right away
        uint32_t pk[16];
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
          pk[c >> 1] = pack_bf16x2(a0, a1);
          pk[(c >> 1) + 1] = pack_bf16x2(b0, b1);
        }
        ptx::tmem_st_32x32b_x16(tmem_s + h * 16, pk);
      }
      float s0, s1, s2, s3;
      ptx::unpack_f32x2(sum2a, s0, s1);
      ptx::unpack_f32x2(sum2b, s2, s3);
      l += (s0 + s1) + (s2 + s3);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[t]);
    }

    // ---- epilogue: O_t / l -> bf16 -> global
    ptx::mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const int row = q0 + t * kBQ + r;
#pragma unroll 1
    for (int c = 0; c < kHD / 32; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < p.Lq) {
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

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace v3

// called by mc_attn_fwd (attn_tcgen05.cu) when the 256-row variant is selected
int32_t launch_attn_v3(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* vt, int64_t ldvt, void* out, int64_t ldo,
                       int32_t Lq, int32_t Lk, int32_t heads, float scale, cudaStream_t stream) {
  const int64_t width = static_cast<int64_t>(heads) * v3::kHD;
  CUtensorMap tq, tk, tv;
  int32_t rc = make_tmap_bf16_2d(&tq, q, static_cast<uint64_t>(Lq), static_cast<uint64_t>(width), static_cast<uint64_t>(ldq), v3::kBQ, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, k, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldk), v3::kBKV, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, vt, static_cast<uint64_t>(width), static_cast<uint64_t>(Lk), static_cast<uint64_t>(ldvt), v3::kHD, 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(v3::attn_fwd_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::kSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn v3 smem)");
    attr_set = true;
  }
  v3::Params p{Lq, Lk, heads, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out), ldo};
  dim3 grid((Lq + 2 * v3::kBQ - 1) / (2 * v3::kBQ), heads);
  v3::attn_fwd_kernel_v3<<<grid, v3::kThreads, v3::kSmem, stream>>>(tq, tk, tv, p);
  MC_CHECK_LAUNCH("attn_fwd_kernel_v3 launch");
  return MC_OK;
}

}  // namespace mc
