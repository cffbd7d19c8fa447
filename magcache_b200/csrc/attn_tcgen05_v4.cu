// Flash-attention forward, split-row softmax: the v1 kernel (attn_tcgen05.cu: 128-row query tile, 64-row KV tiles, P in
// TMEM, two CTAs per SM) with EIGHT softmax warps per CTA instead of four. Two threads share a query row — warp w and
// warp w+4 own the same 32 TMEM lanes and take the left / right 32 columns of every S tile — so each scheduler holds four
// softmax warps (two per co-resident CTA) instead of two, and one warp's latency chain (mbarrier wake-up, tcgen05.ld,
// max, tcgen05.st + wait, fence, arrive) is covered by three others' exponentials rather than one.
// The row max is exchanged between the two halves through 1 KB of shared memory and a 64-thread named barrier per warp
// pair; the row sum stays split until the epilogue. Everything else (operand layouts, pipelines, lazy O rescale) is v1's.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {
namespace v4 {

constexpr int kBQ = 128, kBKV = 64, kHD = 128;
constexpr int kQBytes = kBQ * kHD * 2;   // 32 KB
constexpr int kKBytes = kBKV * kHD * 2;  // 16 KB
constexpr int kVBytes = kHD * kBKV * 2;  // 16 KB
constexpr int kStages = 2;
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kStages * kKBytes;
constexpr int kOffX = kOffV + kStages * kVBytes;  // 96 KB: max exchange [2 buffers][2 halves][128 rows] + sum exchange [2][128], fp32
constexpr int kOffBar = kOffX + 3072;
constexpr int kSmem = kOffBar + 256;
constexpr int kThreads = 320;  // 8 softmax warps + TMA warp + MMA warp
constexpr int kTmemCols = 256; // S0 [0,64) S1 [64,128) O [128,256)
constexpr float kRescaleThreshold = 8.0f;

struct Params {
  int Lq, Lk, heads;
  float scale_log2;
  __nv_bfloat16* out;
  int64_t ldo;
};

__device__ __forceinline__ void pair_sync(int quarter) {  // the two warps that share TMEM lane quarter `quarter`
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
    attn_fwd_kernel_v4(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_vt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* xch = reinterpret_cast<float*>(smem + kOffX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* p_full = bars + 11;  // 256 arrivals
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ;
  const int head = blockIdx.y;
  const int n_tiles = (p.Lk + kBKV - 1) / kBKV;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("attn_fwd_kernel_v4: dynamic smem base not 1024-aligned\n");
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
    }
    ptx::mbar_init(p_full, 256);
    ptx::mbar_init(pv_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 9) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 8) {
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
        ptx::tma_load_2d(smem + kOffK + s * kKBytes, &tmap_k, &k_full[s], head * kHD, j * kBKV);
        ptx::tma_load_2d(smem + kOffK + s * kKBytes + kKBytes / 2, &tmap_k, &k_full[s], head * kHD + 64, j * kBKV);
        ptx::mbar_wait(&v_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&v_full[s], kVBytes);
        ptx::tma_load_2d(smem + kOffV + s * kVBytes, &tmap_vt, &v_full[s], j * kBKV, head * kHD);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(kBQ, kBKV);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(kBQ, kHD);
      const uint32_t q_addr = ptx::smem_u32(smem + kOffQ);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint32_t k_addr = ptx::smem_u32(smem + kOffK + s * kKBytes);
#pragma unroll
        for (int kk = 0; kk < kHD / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_sw128_kmajor(q_addr + (kk >> 2) * (kQBytes / 2)) + 2 * (kk & 3);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(k_addr + (kk >> 2) * (kKBytes / 2)) + 2 * (kk & 3);
          ptx::umma_ss(tmem_base + s * kBKV, da, db, idesc_s, kk != 0 ? 1u : 0u);
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
          ptx::tc_fence_after();
          issue_s(t);  // S buffer t&1 held S/P of tile t-2: ordered behind PV(t-2) by the tensor pipe
        }
        ptx::mbar_wait(p_full, j & 1);
        ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(smem + kOffV + (j & 1) * kVBytes);
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          const uint64_t db = ptx::umma_desc_sw128_kmajor(v_addr) + 2 * kk;
          ptx::umma_ts(tmem_o, tmem_base + (j & 1) * kBKV + kk * 8, db, idesc_o, (j | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&v_empty[j & 1]);
        ptx::umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------ softmax warps (two threads per query row) -------------------
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;  // row inside the Q tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const uint64_t scale2 = ptx::pack_f32x2(p.scale_log2, p.scale_log2);
    float m = -INFINITY, l = 0.f;  // m: row max (identical in both halves); l: this half's share of the row sum

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      ptx::mbar_wait(&s_full[b], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sreg[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_sel + b * kBKV + half * 32, sreg);
      ptx::tmem_ld_wait();
      const int valid = p.Lk - j * kBKV - half * 32;  // columns >= valid of THIS half are padding
      if (valid < 32) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c >= valid) sreg[c] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        mx0 = ptx::max3(mx0, __uint_as_float(sreg[c]), __uint_as_float(sreg[c + 1]));
        mx1 = ptx::max3(mx1, __uint_as_float(sreg[c + 2]), __uint_as_float(sreg[c + 3]));
      }
      float mx = fmaxf(mx0, mx1);
      float* xb = xch + (j & 1) * 256;  // double-buffered: the partner reads tile j's slot before it can reach tile j+1's barrier
      xb[half * 128 + r] = mx;
      pair_sync(quarter);
      mx = fmaxf(mx, xb[(half ^ 1) * 128 + r]);
      const float m_new = fmaxf(m, mx * p.scale_log2);
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = m_new > m + kRescaleThreshold;  // same rows, same values in both halves -> same vote
        if (__any_sync(0xffffffffu, need)) {
          ptx::mbar_wait(pv_done, (j - 1) & 1);  // O quiescent
          ptx::tc_fence_after();
          const float factor = need ? ptx::ex2_approx(m - m_new) : 1.0f;
          if (need) {
            l *= factor;
            m = m_new;
          }
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {  // this half rescales its 64 columns of O, 16 at a time (register budget)
            uint32_t o[16];
            const uint32_t addr = tmem_o + lane_sel + half * 64 + c * 16;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]),
                  "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
                : "r"(addr)
                : "memory");
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            ptx::tmem_st_32x32b_x16(addr, o);
          }
          ptx::tmem_st_wait();
        }
      }
      const uint64_t negm2 = ptx::pack_f32x2(-m, -m);
      uint64_t sum2a = 0ull, sum2b = 0ull;
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        float a0, a1, b0, b1;
        ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[c]), __uint_as_float(sreg[c + 1])), scale2, negm2), a0, a1);
        ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(sreg[c + 2]), __uint_as_float(sreg[c + 3])), scale2, negm2), b0, b1);
        a0 = ptx::ex2_approx(a0);
        a1 = ptx::ex2_approx(a1);
        b0 = ptx::ex2_approx(b0);
        b1 = ptx::ex2_approx(b1);
        sum2a = ptx::add_f32x2(sum2a, ptx::pack_f32x2(a0, a1));
        sum2b = ptx::add_f32x2(sum2b, ptx::pack_f32x2(b0, b1));
        pk[c >> 1] = pack_bf16x2(a0, a1);
        pk[(c >> 1) + 1] = pack_bf16x2(b0, b1);
      }
      float s0, s1, s2, s3;
      ptx::unpack_f32x2(sum2a, s0, s1);
      ptx::unpack_f32x2(sum2b, s2, s3);
      l += (s0 + s1) + (s2 + s3);
      // P_j: 64 bf16 per row = 32 packed words over the first 32 columns of S buffer b; this half owns words [16*half, +16).
      // Word w of the LEFT half's region overlaps S columns the RIGHT half may still be loading only if w >= 32 — it is not.
      ptx::tmem_st_32x32b_x16(tmem_base + lane_sel + b * kBKV + half * 16, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }

    // ---- epilogue: O / (l_left + l_right) -> bf16 -> global; each half writes its 64 columns
    float* xl = xch + 512;  // its own slots: never aliases a max slot a slower partner might still be reading
    xl[half * 128 + r] = l;
    pair_sync(quarter);
    const float inv_l = 1.0f / (l + xl[(half ^ 1) * 128 + r]);
    ptx::mbar_wait(pv_done, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
    const int row = q0 + r;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + lane_sel + half * 64 + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < p.Lq) {
        __nv_bfloat16* dst = p.out + static_cast<int64_t>(row) * p.ldo + head * kHD + half * 64 + c * 32;
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

}  // namespace v4

int32_t launch_attn_v4(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* vt, int64_t ldvt, void* out, int64_t ldo,
                       int32_t Lq, int32_t Lk, int32_t heads, float scale, cudaStream_t stream) {
  const int64_t width = static_cast<int64_t>(heads) * v4::kHD;
  CUtensorMap tq, tk, tv;
  int32_t rc = make_tmap_bf16_2d(&tq, q, static_cast<uint64_t>(Lq), static_cast<uint64_t>(width), static_cast<uint64_t>(ldq), v4::kBQ, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, k, static_cast<uint64_t>(Lk), static_cast<uint64_t>(width), static_cast<uint64_t>(ldk), v4::kBKV, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, vt, static_cast<uint64_t>(width), static_cast<uint64_t>(Lk), static_cast<uint64_t>(ldvt), v4::kHD, v4::kBKV);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(v4::attn_fwd_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, v4::kSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn v4 smem)");
    attr_set = true;
  }
  v4::Params p{Lq, Lk, heads, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(out), ldo};
  dim3 grid((Lq + v4::kBQ - 1) / v4::kBQ, heads);
  v4::attn_fwd_kernel_v4<<<grid, v4::kThreads, v4::kSmem, stream>>>(tq, tk, tv, p);
  MC_CHECK_LAUNCH("attn_fwd_kernel_v4 launch");
  return MC_OK;
}

}  // namespace mc
