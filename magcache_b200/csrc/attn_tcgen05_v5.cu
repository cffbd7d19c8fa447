// Flash-attention forward, 128-wide KV tiles with two CTAs per SM. The round-1 profile of the 64-wide kernel
// (profiles/r01_ncu_attn.md) shows the tensor-core pipe 87 % occupied for 68 % of MMA math: the S = Q K^T MMA at N = 64
// re-reads the 32 KB Q tile from shared memory for every 16 KB K tile (192 B/clk of operands against a 128 B/clk port).
// At N = 128 the same Q read feeds twice the MACs (64 KB of operands per 512 math cycles = the port rate), and there are half
// as many softmax round trips per key. TMEM stays at 256 columns per CTA (two CTAs per SM) by single-buffering S:
//   S [0,128) (P, 64 packed bf16 columns, overwrites the consumed S) | O [128,256)
// and shared memory stays under 113 KB with single-stage K and V^T tiles (each is re-fillable long before it is needed:
// K after the S MMA, V^T after the PV MMA, both ~2000 cycles ahead of their next use). Inside one CTA the chain
// S -> softmax -> PV is serial; the two co-resident CTAs fill each other's gaps.
//   warps 0-3 softmax (thread = query row), warp 4 TMA, warp 5 MMA
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {
namespace v5 {

constexpr int kBQ = 128, kBKV = 128, kHD = 128;
constexpr int kQTileBytes = kBQ * kHD * 2;   // 32 KB (two 64-column boxes)
constexpr int kKBytes = kBKV * kHD * 2;      // 32 KB (two boxes [128 kv x 64 hd])
constexpr int kVBytes = kHD * kBKV * 2;      // 32 KB (two boxes [128 d x 64 kv])
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQTileBytes;   // 32 KB
constexpr int kOffV = kOffK + kKBytes;       // 64 KB
constexpr int kOffBar = kOffV + kVBytes;     // 96 KB
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct Params {
  int Lq, Lk, heads;
  float scale_log2;
  __nv_bfloat16* out;
  int64_t ldo;
};

__global__ void __launch_bounds__(kThreads, 2)
    attn_fwd_kernel_v5(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_vt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;   // 128 arrivals
  uint64_t* pv_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ;
  const int head = blockIdx.y;
  const int n_tiles = (p.Lk + kBKV - 1) / kBKV;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_fwd_kernel_v5: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_vt);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(k_full, 1);
    ptx::mbar_init(k_empty, 1);
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(pv_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 5) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, kQTileBytes);
      ptx::tma_load_2d(smem + kOffQ, &tmap_q, q_full, head * kHD, q0);
      ptx::tma_load_2d(smem + kOffQ + kQTileBytes / 2, &tmap_q, q_full, head * kHD + 64, q0);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        ptx::mbar_wait(k_empty, ph ^ 1);  // S(j-1) has read the K buffer
        ptx::mbar_expect_tx(k_full, kKBytes);
        ptx::tma_load_2d(smem + kOffK, &tmap_k, k_full, head * kHD, j * kBKV);
        ptx::tma_load_2d(smem + kOffK + kKBytes / 2, &tmap_k, k_full, head * kHD + 64, j * kBKV);
        ptx::mbar_wait(v_empty, ph ^ 1);  // PV(j-1) has read the V^T buffer
        ptx::mbar_expect_tx(v_full, kVBytes);
        ptx::tma_load_2d(smem + kOffV, &tmap_vt, v_full, j * kBKV, head * kHD);
        ptx::tma_load_2d(smem + kOffV + kVBytes / 2, &tmap_vt, v_full, j * kBKV + 64, head * kHD);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kBQ, 128);  // both MMAs are 128 x 128 (x K16)
      const uint32_t q_addr = ptx::smem_u32(smem + kOffQ);
      const uint32_t k_addr = ptx::smem_u32(smem + kOffK);
      const uint32_t v_addr = ptx::smem_u32(smem + kOffV);
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        // S(j) overwrites S/P(j-1): PV(j-1) was issued before it (in-order tensor pipe) and p_full(j-1) was waited for
        ptx::mbar_wait(k_full, j & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kHD / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_sw128_kmajor(q_addr + (kk >> 2) * (kQTileBytes / 2)) + 2 * (kk & 3);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(k_addr + (kk >> 2) * (kKBytes / 2)) + 2 * (kk & 3);
          ptx::umma_ss(tmem_base, da, db, idesc, kk != 0 ? 1u : 0u);
        }
        ptx::umma_commit(k_empty);
        ptx::umma_commit(s_full);
        ptx::mbar_wait(p_full, j & 1);
        ptx::mbar_wait(v_full, j & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {  // O += P(j) V_j : A from TMEM (8 packed columns per K16 step)
          const uint64_t db = ptx::umma_desc_sw128_kmajor(v_addr + (kk >> 2) * (kVBytes / 2)) + 2 * (kk & 3);
          ptx::umma_ts(tmem_base + 128, tmem_base + kk * 8, db, idesc, (j | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(v_empty);
        ptx::umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------ softmax warpgroups ------------------------------------------
    const int r = warp * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_sel;
    const uint32_t tmem_o = tmem_base + 128 + lane_sel;
    const uint64_t scale2 = ptx::pack_f32x2(p.scale_log2, p.scale_log2);
    float m = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(s_full, j & 1);
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
          ptx::mbar_wait(pv_done, (j - 1) & 1);  // O quiescent
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
      ptx::mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> global
    ptx::mbar_wait(pv_done, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const int row = q0 + r;
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
  if (warp == 5) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace v5

// called by mc_attn_fwd (attn_tcgen05.cu) when MC_ATTN_VARIANT=5
int32_t launch_attn_v5(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* vt, int64_t ldvt, void* out, int64_t ldo,
                       int32_t Lq, int32_t Lk, int32_t heads, float scale, cudaStream_t stream) {
  const int64_t width = static_cast<int64_t>(heads) * v5::kHD;
  CUtensorMap tq, tk, tv;
  int32_t rc = make_tmap_bf16_2d(&tq, q, static_cast<uint64_t>(Lq), static_cast<uint64_t>(width), static_cast<uint64_t>(ldq), v5::kBQ, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, k, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldk), v5::kBKV, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, vt, static_cast<uint64_t>(width), static_cast<uint64_t>(Lk), static_cast<uint64_t>(ldvt), v5::kHD, 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(v5::attn_fwd_kernel_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, v5::kSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn v5 smem)");
    attr_set = true;
  }
  v5::Params p{Lq, Lk, heads, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out), ldo};
  dim3 grid((Lq + v5::kBQ - 1) / v5::kBQ, heads);
  v5::attn_fwd_kernel_v5<<<grid, v5::kThreads, v5::kSmem, stream>>>(tq, tk, tv, p);
  MC_CHECK_LAUNCH("attn_fwd_kernel_v5 launch");
  return MC_OK;
}

}  // namespace mc
