// bf16 GEMM on the 5th-gen tensor cores: acc[m,n] = sum_k A[m,k] * B[n,k], fp32 accumulation in TMEM.
// Replaces the cuBLAS calls behind every nn.Linear / Conv3d of the cache-miss branch
// (`for block in self.blocks: x = block(x, **kwargs)`, MagCache4Wan2.1/magcache_generate.py:297-298; patch_embedding :237;
// text_embedding :257-262), with the reference's elementwise follow-ups fused into the epilogue.
//
// Tile: 128 (M) x 128 (N) x 64 (K) per pipeline stage, one output tile per CTA, two CTAs co-resident per SM so one CTA's
// epilogue overlaps the other's main loop. Warp roles (192 threads):
//   warp 0   TMA producer  : cp.async.bulk.tensor 2-D loads of A and B tiles (128-byte swizzle) into a 3-stage smem ring
//   warp 1   MMA issuer    : one thread issues tcgen05.mma (M128 N128 K16) x4 per stage; tcgen05.commit frees the stage
//   warps 2-5 epilogue     : tcgen05.ld the 128x128 fp32 accumulator (lane = row), apply the epilogue, store to global
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mc {

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 3;
constexpr int kTileABytes = kBM * kBK * 2;  // 16 KB
constexpr int kTileBBytes = kBN * kBK * 2;  // 16 KB
constexpr int kStageBytes = kTileABytes + kTileBBytes;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /* alignment slack */ + 128 /* barriers */;
constexpr int kGemmThreads = 192;

struct GemmParams {
  int M, N, K;
  const float* bias;
  void* out;
  int64_t ldo;
  const float* gate;
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, const uint32_t (&acc)[32]) {
  // `row` is in range; columns col0 .. col0+31 may run past N.
  const int ncols = min(32, p.N - col0);
  if (ncols <= 0) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);

  if (EPI == MC_EPI_ROWBIAS_BF16) {
    const float b = p.bias ? p.bias[row] : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += b;
  } else {
    if (p.bias) {
      if (ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + j);
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      } else {
        for (int j = 0; j < ncols; ++j) v[j] += p.bias[col0 + j];
      }
    }
  }

  if (EPI == MC_EPI_BIAS_F32) {
    float* o = static_cast<float*>(p.out) + static_cast<int64_t>(row) * p.ldo + col0;
    if (ncols == 32 && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < ncols; ++j) o[j] = v[j];
    }
    return;
  }

  if (EPI == MC_EPI_BIAS_GATE_RESID) {
    // x[m,n] = x[m,n] + float(bf16(acc + bias)) * gate[n]     (`x = x + y * e[2]` in fp32, y is the bf16 Linear output)
    float* o = static_cast<float*>(p.out) + static_cast<int64_t>(row) * p.ldo + col0;
    if (ncols == 32 && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 x = *reinterpret_cast<const float4*>(o + j);
        float g0 = 1.f, g1 = 1.f, g2 = 1.f, g3 = 1.f;
        if (p.gate) {
          const float4 g = *reinterpret_cast<const float4*>(p.gate + col0 + j);
          g0 = g.x; g1 = g.y; g2 = g.z; g3 = g.w;
        }
        x.x = __fadd_rn(x.x, __fmul_rn(round_bf16(v[j]), g0));
        x.y = __fadd_rn(x.y, __fmul_rn(round_bf16(v[j + 1]), g1));
        x.z = __fadd_rn(x.z, __fmul_rn(round_bf16(v[j + 2]), g2));
        x.w = __fadd_rn(x.w, __fmul_rn(round_bf16(v[j + 3]), g3));
        *reinterpret_cast<float4*>(o + j) = x;
      }
    } else {
      for (int j = 0; j < ncols; ++j) {
        const float g = p.gate ? p.gate[col0 + j] : 1.f;
        o[j] = __fadd_rn(o[j], __fmul_rn(round_bf16(v[j]), g));
      }
    }
    return;
  }

  // bf16 outputs
  if (EPI == MC_EPI_BIAS_GELU_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(round_bf16(v[j]));  // Linear output is bf16 before nn.GELU sees it
  }
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(row) * p.ldo + col0;
  if (ncols == 32 && (p.ldo & 7) == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 w;
      w.x = pack_bf16x2(v[j], v[j + 1]);
      w.y = pack_bf16x2(v[j + 2], v[j + 3]);
      w.z = pack_bf16x2(v[j + 4], v[j + 5]);
      w.w = pack_bf16x2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(o + j) = w;
    }
  } else {
    for (int j = 0; j < ncols; ++j) o[j] = __float2bfloat16_rn(v[j]);
  }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 2)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // N tiles vary fastest so the CTAs that share an A row-panel run together (A panel stays in L2; B is small)
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int m0 = (blockIdx.x / n_tiles) * kBM;
  const int n0 = (blockIdx.x % n_tiles) * kBN;
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kBN);  // 128 fp32 accumulator columns
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kTileABytes;
        ptx::mbar_expect_tx(&full_bar[s], kStageBytes);
        ptx::tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBK, m0);
        ptx::tma_load_2d(sb, &tmap_b, &full_bar[s], kb * kBK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kBM, kBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * kStageBytes);
        const uint64_t da = ptx::umma_desc_sw128_kmajor(sa);
        const uint64_t db = ptx::umma_desc_sw128_kmajor(sa + kTileABytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzled row: +2 in the (addr >> 4) field
          ptx::umma_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
      }
      ptx::umma_commit(acc_bar);  // accumulator complete
    }
  } else {
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t acc[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, acc);
      ptx::tmem_ld_wait();
      if (row < p.M) epilogue_chunk<EPI>(p, row, n0 + c * 32, acc);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kBN);
}

template <int EPI>
static int32_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm smem)");
    attr_set = true;
  }
  const int m_tiles = (p.M + kBM - 1) / kBM, n_tiles = (p.N + kBN - 1) / kBN;
  gemm_bf16_kernel<EPI><<<m_tiles * n_tiles, kGemmThreads, kGemmSmem, s>>>(ta, tb, p);
  MC_CHECK_LAUNCH("gemm_bf16_kernel launch");
  return MC_OK;
}

}  // namespace mc

extern "C" int32_t mc_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t M, int32_t N, int32_t K,
                                const float* bias, int32_t epilogue, void* out, int64_t ldo, const float* gate, void* stream) {
  MC_CHECK_ARG(A && B && out, "mc_gemm_bf16: null pointer");
  MC_CHECK_ARG(M >= 1 && N >= 1 && K >= 8, "mc_gemm_bf16: M=%d N=%d K=%d", M, N, K);
  MC_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, "mc_gemm_bf16: K/lda/ldb must be multiples of 8 (16-byte TMA pitch)");
  MC_CHECK_ARG(mc::aligned16(A) && mc::aligned16(B) && mc::aligned16(out), "mc_gemm_bf16: A/B/out must be 16-byte aligned");
  MC_CHECK_ARG(ldo >= N, "mc_gemm_bf16: ldo=%lld < N=%d", static_cast<long long>(ldo), N);
  MC_CHECK_ARG(bias == nullptr || mc::aligned16(bias), "mc_gemm_bf16: bias must be 16-byte aligned");
  MC_CHECK_ARG(gate == nullptr || mc::aligned16(gate), "mc_gemm_bf16: gate must be 16-byte aligned");
  CUtensorMap ta, tb;
  int32_t rc = mc::make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), mc::kBM, mc::kBK);
  if (rc) return rc;
  rc = mc::make_tmap_bf16_2d(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), mc::kBN, mc::kBK);
  if (rc) return rc;
  mc::GemmParams p{M, N, K, bias, out, ldo, gate};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case MC_EPI_BIAS_BF16: return mc::launch_gemm<MC_EPI_BIAS_BF16>(ta, tb, p, s);
    case MC_EPI_BIAS_GELU_BF16: return mc::launch_gemm<MC_EPI_BIAS_GELU_BF16>(ta, tb, p, s);
    case MC_EPI_BIAS_GATE_RESID: return mc::launch_gemm<MC_EPI_BIAS_GATE_RESID>(ta, tb, p, s);
    case MC_EPI_ROWBIAS_BF16: return mc::launch_gemm<MC_EPI_ROWBIAS_BF16>(ta, tb, p, s);
    case MC_EPI_BIAS_F32: return mc::launch_gemm<MC_EPI_BIAS_F32>(ta, tb, p, s);
    default:
      mc::set_error("mc_gemm_bf16: unknown epilogue %d", epilogue);
      return MC_ERR_INVALID;
  }
}
