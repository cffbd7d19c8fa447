// bf16 GEMM on the 5th-gen tensor cores: acc[m,n] = sum_k A[m,k] * B[n,k], fp32 accumulation in TMEM.
// Replaces the cuBLAS calls behind every nn.Linear / Conv3d of the cache-miss branch
// (`for block in self.blocks: x = block(x, **kwargs)`, MagCache4Wan2.1/magcache_generate.py:297-298; patch_embedding :237;
// text_embedding :257-262), with the reference's elementwise follow-ups fused into the epilogue.
//
// Persistent kernel, one CTA per SM, tile 128 (M) x 256 (N) x 64 (K per stage):
//   warp 0     TMA producer : A tile (128x64) + B tile (256x64) per stage, 128-byte swizzle, 4-stage smem ring (192 KB)
//   warp 1     MMA issuer   : tcgen05.mma M128 N256 K16 x4 per stage into one of TWO 256-column TMEM accumulators
//   warps 2-9  epilogue     : drain accumulator t&1 while the MMA warp fills the other one; two warps per TMEM lane quarter,
//                             each taking 128 of the 256 columns. tcgen05.ld (lane = row) -> per-warp smem transpose ->
//                             lane = column, so every global access is a coalesced row segment. (With four epilogue warps
//                             the K = 1536 GEMMs were epilogue-bound: ~12 us of dependent load/store per tile against a
//                             6.4 us main loop.)
// Tiles are walked in waves of gridDim.x consecutive tiles with N fastest: the CTAs of one wave share A row-panels in L2.
// A second instantiation with 128-wide N tiles serves small grids (the text K|V projection: 512 x 3072 is 48 wide tiles on 148
// SMs, 96 narrow ones: 21 -> 18 us). A narrow tile is NOT half the cost of a wide one — the same A panel moves for half the math and
// the operands sit at the 128 B/clk fetch limit: measured 0.62-0.65 — so for the ragged grids of a token-sharded rank (4095 x 1536:
// two wide waves at 65 % against three narrow ones) it gains 6 % at K = 1536 and LOSES 18 % at K = 8960; pick_bn's cost model
// therefore only takes it where the wave count makes it at least 8 % cheaper at 0.62 per wave.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {

TmapCacheEntry* tmap_cache() {
  static TmapCacheEntry table[kTmapCacheSize] = {};
  return table;
}
std::mutex& tmap_cache_mutex() {
  static std::mutex m;
  return m;
}

constexpr int kBM = 128, kBK = 64, kStages = 4;
constexpr int kTileABytes = kBM * kBK * 2;  // 16 KB
constexpr int kStagePad = 33;                                 // floats per staged row (conflict-free transpose)
constexpr int kEpiWarps = 8;
constexpr int kStagingBytes = kEpiWarps * 32 * kStagePad * 4;  // one 32x32 fp32 patch per epilogue warp
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + epilogue warps
template <int BN>
struct GemmTile {  // BN = 256 (default) or 128
  static constexpr int kTileBBytes = BN * kBK * 2;                 // 32 / 16 KB
  static constexpr int kStageBytes = kTileABytes + kTileBBytes;
  static constexpr int kOffStaging = kStages * kStageBytes;        // 196608 / 131072
  static constexpr int kOffBars = kOffStaging + kStagingBytes;     // + 33792
  static constexpr int kSmem = kOffBars + 128;
  static constexpr int kTmemCols = 2 * BN;                         // 2 accumulators x BN fp32 columns
};

struct GemmParams {
  int M, N, K;
  const float* bias;
  void* out;
  int64_t ldo;
  const float* gate;
};

// One 32-row x 32-column patch: `stage` holds acc[r][c] at stage[r*33 + c]; lane = column. All global accesses below
// touch 128 (fp32) or 64 (bf16, two rows per instruction) contiguous bytes per row.
template <int EPI>
__device__ __forceinline__ void epilogue_patch(const GemmParams& p, const float* stage, int row0, int col0, int lane) {
  const int rows = min(32, p.M - row0);
  if (rows <= 0) return;
  const int col = col0 + lane;
  const bool col_ok = col < p.N;

  if (EPI == MC_EPI_BIAS_GATE_RESID) {
    // x[m,n] = x[m,n] + float(bf16(acc + bias[n])) * gate[n]   (`x = x + y * e[2]`: y is the bf16 Linear output, fp32 stream)
    const float b = (p.bias && col_ok) ? p.bias[col] : 0.f;
    const float g = (p.gate && col_ok) ? p.gate[col] : 1.f;
    float* xcol = static_cast<float*>(p.out) + static_cast<int64_t>(row0) * p.ldo + col;
    float xv[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) xv[r] = (col_ok && r < rows) ? xcol[static_cast<int64_t>(r) * p.ldo] : 0.f;  // 32 loads in flight
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float y = round_bf16(stage[r * kStagePad + lane] + b);
      xv[r] = __fadd_rn(xv[r], __fmul_rn(y, g));
    }
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (col_ok && r < rows) xcol[static_cast<int64_t>(r) * p.ldo] = xv[r];
    return;
  }

  if (EPI == MC_EPI_BIAS_F32) {
    const float b = (p.bias && col_ok) ? p.bias[col] : 0.f;
    float* ocol = static_cast<float*>(p.out) + static_cast<int64_t>(row0) * p.ldo + col;
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (col_ok && r < rows) ocol[static_cast<int64_t>(r) * p.ldo] = stage[r * kStagePad + lane] + b;
    return;
  }

  // bf16 outputs: each instruction writes two rows x 16 column pairs (lane -> row parity = lane/16, column pair = lane%16)
  const int half = lane >> 4, cp = (lane & 15) * 2;
  const int c0 = col0 + cp;
  float b0 = 0.f, b1 = 0.f;
  if (EPI != MC_EPI_ROWBIAS_BF16 && p.bias) {
    if (c0 < p.N) b0 = p.bias[c0];
    if (c0 + 1 < p.N) b1 = p.bias[c0 + 1];
  }
  float g0 = 1.f, g1 = 1.f;
  if (EPI == MC_EPI_BIAS_GATE_RESID_BF16 && p.gate) {
    if (c0 < p.N) g0 = p.gate[c0];
    if (c0 + 1 < p.N) g1 = p.gate[c0 + 1];
  }
  __nv_bfloat16* obase = static_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(row0 + half) * p.ldo + c0;
  const float* sbase = stage + half * kStagePad + cp;
  const float* rbias = (EPI == MC_EPI_ROWBIAS_BF16 && p.bias) ? p.bias + row0 + half : nullptr;

  if (rows == 32 && col0 + 32 <= p.N && (p.ldo & 1) == 0) {
    // interior patch: branch-free, 16 independent iterations for the scheduler to interleave
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v0 = sbase[2 * i * kStagePad], v1 = sbase[2 * i * kStagePad + 1];
      if (EPI == MC_EPI_ROWBIAS_BF16) {
        const float rb = rbias ? rbias[2 * i] : 0.f;  // V^T = Wv h^T + bv: bias indexed by the output ROW
        v0 += rb;
        v1 += rb;
      } else {
        v0 += b0;
        v1 += b1;
      }
      if (EPI == MC_EPI_BIAS_GELU_BF16) {  // the Linear output is bf16 before nn.GELU(tanh) sees it
        v0 = gelu_tanh(round_bf16(v0));
        v1 = gelu_tanh(round_bf16(v1));
      }
      if (EPI == MC_EPI_BIAS_GELU_ERF_BF16) {
        v0 = gelu_erf(round_bf16(v0));
        v1 = gelu_erf(round_bf16(v1));
      }
      if (EPI == MC_EPI_BIAS_SILU_BF16) {
        v0 = silu_f(round_bf16(v0));
        v1 = silu_f(round_bf16(v1));
      }
      if (EPI == MC_EPI_BIAS_GATE_RESID_BF16) {  // x = x + g * y with every tensor bf16 (MMDiT streams): three roundings
        const uint32_t old = *reinterpret_cast<const uint32_t*>(obase + static_cast<int64_t>(2 * i) * p.ldo);
        v0 = bf16_lo(old) + round_bf16(g0 * round_bf16(v0));
        v1 = bf16_hi(old) + round_bf16(g1 * round_bf16(v1));
      }
      w[i] = pack_bf16x2(v0, v1);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(2 * i) * p.ldo) = w[i];
    return;
  }

  const bool pair_ok = (c0 + 1 < p.N) && ((p.ldo & 1) == 0);
  for (int i = 0; i < 16; ++i) {
    const int r = 2 * i + half;
    if (r >= rows) continue;
    float v0 = sbase[2 * i * kStagePad], v1 = sbase[2 * i * kStagePad + 1];
    if (EPI == MC_EPI_ROWBIAS_BF16) {
      const float rb = rbias ? rbias[2 * i] : 0.f;
      v0 += rb;
      v1 += rb;
    } else {
      v0 += b0;
      v1 += b1;
    }
    if (EPI == MC_EPI_BIAS_GELU_BF16) {
      v0 = gelu_tanh(round_bf16(v0));
      v1 = gelu_tanh(round_bf16(v1));
    }
    if (EPI == MC_EPI_BIAS_GELU_ERF_BF16) {
      v0 = gelu_erf(round_bf16(v0));
      v1 = gelu_erf(round_bf16(v1));
    }
    if (EPI == MC_EPI_BIAS_SILU_BF16) {
      v0 = silu_f(round_bf16(v0));
      v1 = silu_f(round_bf16(v1));
    }
    __nv_bfloat16* o = obase + static_cast<int64_t>(2 * i) * p.ldo;
    if (EPI == MC_EPI_BIAS_GATE_RESID_BF16) {
      if (c0 < p.N) v0 = __bfloat162float(o[0]) + round_bf16(g0 * round_bf16(v0));
      if (c0 + 1 < p.N) v1 = __bfloat162float(o[1]) + round_bf16(g1 * round_bf16(v1));
    }
    if (pair_ok) {
      *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v0, v1);
    } else {
      if (c0 < p.N) o[0] = __float2bfloat16_rn(v0);
      if (c0 + 1 < p.N) o[1] = __float2bfloat16_rn(v1);
    }
  }
}

template <int EPI, int kBN>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using T = GemmTile<kBN>;
  constexpr int kStageBytes = T::kStageBytes, kOffStaging = T::kOffStaging, kOffBars = T::kOffBars, kTmemCols = T::kTmemCols;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* staging = reinterpret_cast<float*>(smem + kOffStaging);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + kBN - 1) / kBN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("gemm_bf16_kernel: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 32 * kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: the whole warp walks the loop in uniform control flow, one elected lane issues the copies (ptx::elect_one)
    uint32_t it = 0;  // running k-block counter across tiles -> smem stage / phase
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        if (ptx::elect_one()) {
          uint8_t* sa = smem + s * kStageBytes;
          ptx::mbar_expect_tx(&full_bar[s], kStageBytes);
          ptx::tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBK, m0);
          ptx::tma_load_2d(sa + kTileABytes, &tmap_b, &full_bar[s], kb * kBK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // MMA issuer: all lanes wait on the barriers, the elected lane issues the four K16 MMAs of a stage and the commits
    constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kBM, kBN);
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1;
      ptx::mbar_wait(&acc_empty[acc], ((t >> 1) & 1) ^ 1);  // epilogue has drained this accumulator (passes on first use)
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kBN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + s * kStageBytes);
          const uint64_t da = ptx::umma_desc_sw128_kmajor(sa);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(sa + kTileABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzled row: +2 in the (addr >> 4) field
            ptx::umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs have read it
          if (kb == num_kb - 1) ptx::umma_commit(&acc_full[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;           // TMEM lane quarter this warp may access (warps 2..9 -> 2,3,0,1,2,3,0,1)
    const int chalf = (warp - 2) >> 2;  // which 128-column half of the tile this warp drains
    float* stage = staging + (warp - 2) * 32 * kStagePad;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
      const int m0 = (tile / n_tiles) * kBM, n0 = (tile % n_tiles) * kBN;
      const uint32_t acc = t & 1;
      ptx::mbar_wait(&acc_full[acc], (t >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kBN + (static_cast<uint32_t>(q * 32) << 16);
      const int row0 = m0 + q * 32;
#pragma unroll 1
      for (int c = chalf * (kBN / 64); c < (chalf + 1) * (kBN / 64); ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
        ptx::tmem_ld_wait();
        if (c == (chalf + 1) * (kBN / 64) - 1) {  // this warp's last read of the accumulator: hand it back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_empty[acc]);
        }
        __syncwarp();  // previous patch fully consumed before the staging rows are overwritten
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * kStagePad + j] = __uint_as_float(v[j]);
        __syncwarp();
        epilogue_patch<EPI>(p, stage, row0, n0 + c * 32, lane);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

template <int EPI, int BN>
static int32_t launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  static PerDeviceOnce once;
  const int32_t rc = set_max_smem_once(gemm_bf16_kernel<EPI, BN>, GemmTile<BN>::kSmem, once, "cudaFuncSetAttribute(gemm smem)");
  if (rc) return rc;
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_bf16_kernel<EPI, BN><<<grid, kGemmThreads, GemmTile<BN>::kSmem, s>>>(ta, tb, p);
  MC_CHECK_LAUNCH("gemm_bf16_kernel launch");
  return MC_OK;
}

// N-tile width for an M x N output: waves of 148 tiles, a wave of 128-wide tiles at 0.62 of a wave of 256-wide ones (measured, see
// the header). MC_GEMM_BN = 128 | 256 forces one (tests, A/B timing).
static int pick_bn(int M, int N) {
  const char* e = getenv("MC_GEMM_BN");
  if (e && *e) {
    const int v = atoi(e);
    if (v == 128 || v == 256) return v;
  }
  const int64_t sms = num_sms(), m_tiles = (M + kBM - 1) / kBM;
  const int64_t t256 = m_tiles * ((N + 255) / 256), t128 = m_tiles * ((N + 127) / 128);
  const int64_t cost256 = ((t256 + sms - 1) / sms) * 100, cost128 = ((t128 + sms - 1) / sms) * 62;
  return cost128 * 100 <= cost256 * 92 ? 128 : 256;
}

}  // namespace mc

extern "C" int32_t mc_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int32_t K,
                                const float* bias, int32_t epilogue, void* out, int64_t ldo, const float* gate, void* stream) {
  MC_CHECK_ARG(A && B && out, "mc_gemm_bf16: null pointer");
  MC_CHECK_ARG(M >= 1 && N >= 1 && K >= 8, "mc_gemm_bf16: M=%d N=%d K=%d", M, N, K);
  MC_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, "mc_gemm_bf16: K/lda/ldb must be multiples of 8 (16-byte TMA pitch)");
  MC_CHECK_ARG(mc::aligned16(A) && mc::aligned16(B) && mc::aligned16(out), "mc_gemm_bf16: A/B/out must be 16-byte aligned");
  MC_CHECK_ARG(ldo >= N, "mc_gemm_bf16: ldo=%lld < N=%d", static_cast<long long>(ldo), N);
  CUtensorMap ta, tb;
  int32_t rc = mc::make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), mc::kBM, mc::kBK);
  if (rc) return rc;
  const int bn = mc::pick_bn(M, N);
  rc = mc::make_tmap_bf16_2d(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), bn, mc::kBK);
  if (rc) return rc;
  mc::GemmParams p{M, N, K, bias, out, ldo, gate};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define MC_GEMM_CASE(E) \
  case E: return bn == 128 ? mc::launch_gemm_bn<E, 128>(ta, tb, p, s) : mc::launch_gemm_bn<E, 256>(ta, tb, p, s)
  switch (epilogue) {
    MC_GEMM_CASE(MC_EPI_BIAS_BF16);
    MC_GEMM_CASE(MC_EPI_BIAS_GELU_BF16);
    MC_GEMM_CASE(MC_EPI_BIAS_GATE_RESID);
    MC_GEMM_CASE(MC_EPI_ROWBIAS_BF16);
    MC_GEMM_CASE(MC_EPI_BIAS_F32);
    MC_GEMM_CASE(MC_EPI_BIAS_GELU_ERF_BF16);
    MC_GEMM_CASE(MC_EPI_BIAS_GATE_RESID_BF16);
    MC_GEMM_CASE(MC_EPI_BIAS_SILU_BF16);
#undef MC_GEMM_CASE
    default:
      mc::set_error("mc_gemm_bf16: unknown epilogue %d", epilogue);
      return MC_ERR_INVALID;
  }
}
