// Residual-cache kernels — the HBM-bound half of the MagCache hot path.
//   K1 cache-hit add     `x = x + residual_x`            MagCache4Wan2.1/magcache_generate.py:295
//   K2 residual sub      `residual_x = x - ori_x`        MagCache4Wan2.1/magcache_generate.py:299
//   K3 calibration stats  norm ratio / std / cos distance MagCache4Wan2.1/magcache_generate.py:167-169
// Layout: flat contiguous tensors; every thread moves 128-bit words with L1::no_allocate (streamed once),
// UNROLL independent word-groups in flight per thread, grid = a multiple of the SM count (persistent grid-stride).
#include "common.cuh"
#include "ptx.cuh"

namespace mc {

// ---- 8-element group load/store for fp32 and bf16 ----------------------------------------------------------
template <int DT>
struct Elem;
template <>
struct Elem<MC_F32> {
  static constexpr int kBytes = 4;
  static constexpr int kAlign = 32;  // 8 fp32 move as ONE 256-bit access per lane
  __device__ static __forceinline__ void load8(const void* base, int64_t i, float (&f)[8]) {
    ptx::ld_nc_v8_f32(static_cast<const float*>(base) + i, f);
  }
  __device__ static __forceinline__ void store8(void* base, int64_t i, const float (&f)[8]) {
    ptx::st_na_v8_f32(static_cast<float*>(base) + i, f);
  }
  __device__ static __forceinline__ float load1(const void* base, int64_t i) { return static_cast<const float*>(base)[i]; }
  __device__ static __forceinline__ void store1(void* base, int64_t i, float v) { static_cast<float*>(base)[i] = v; }
};
template <>
struct Elem<MC_BF16> {
  static constexpr int kBytes = 2;
  static constexpr int kAlign = 16;
  __device__ static __forceinline__ void load8(const void* base, int64_t i, float (&f)[8]) {
    uint4 a = ptx::ld_nc_v4(static_cast<const __nv_bfloat16*>(base) + i);
    unpack_bf16x8(a, f);
  }
  __device__ static __forceinline__ void store8(void* base, int64_t i, const float (&f)[8]) {
    ptx::st_na_v4(static_cast<__nv_bfloat16*>(base) + i, pack_bf16x8(f));
  }
  __device__ static __forceinline__ float load1(const void* base, int64_t i) {
    return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
  }
  __device__ static __forceinline__ void store1(void* base, int64_t i, float v) {
    static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  }
};

// out = a + sign * b  (fp32 arithmetic, one rounding into the output type: torch's promoted add/sub)
template <int DA, int DB, int DO, int UNROLL>
__global__ void __launch_bounds__(256) axpb_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out,
                                                   int64_t n_groups /* of 8 */, float sign) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // main: UNROLL groups per thread per trip, all loads issued before any store
  for (; g + (UNROLL - 1) * stride < n_groups; g += UNROLL * stride) {
    float fa[UNROLL][8], fb[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      Elem<DA>::load8(a, (g + u * stride) * 8, fa[u]);
      Elem<DB>::load8(b, (g + u * stride) * 8, fb[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) fa[u][j] = fa[u][j] + sign * fb[u][j];
      Elem<DO>::store8(out, (g + u * stride) * 8, fa[u]);
    }
  }
  for (; g < n_groups; g += stride) {
    float fa[8], fb[8];
    Elem<DA>::load8(a, g * 8, fa);
    Elem<DB>::load8(b, g * 8, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] = fa[j] + sign * fb[j];
    Elem<DO>::store8(out, g * 8, fa);
  }
}

template <int DA, int DB, int DO>
__global__ void axpb_scalar_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out, int64_t begin,
                                   int64_t n, float sign) {
  int64_t i = begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) Elem<DO>::store1(out, i, Elem<DA>::load1(a, i) + sign * Elem<DB>::load1(b, i));
}

template <int DA, int DB, int DO>
static int32_t launch_axpb(const void* a, const void* b, void* out, int64_t n, float sign, cudaStream_t s) {
  constexpr int kUnroll = 4;
  auto al = [](const void* p, int a_) { return (reinterpret_cast<uintptr_t>(p) & static_cast<uintptr_t>(a_ - 1)) == 0; };
  const bool vec_ok = al(a, Elem<DA>::kAlign) && al(b, Elem<DB>::kAlign) && al(out, Elem<DO>::kAlign);
  const int64_t n_groups = vec_ok ? n / 8 : 0;
  if (n_groups > 0) {
    const int threads = 256;
    int64_t want = (n_groups + static_cast<int64_t>(threads) * kUnroll - 1) / (static_cast<int64_t>(threads) * kUnroll);
    const int64_t cap = static_cast<int64_t>(num_sms()) * 8;  // 8 resident CTAs of 256 threads per SM = full occupancy
    int grid = static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
    axpb_kernel<DA, DB, DO, kUnroll><<<grid, threads, 0, s>>>(a, b, out, n_groups, sign);
    MC_CHECK_LAUNCH("axpb_kernel launch");
  }
  const int64_t done = n_groups * 8;
  if (done < n) {
    const int64_t rem = n - done;
    axpb_scalar_kernel<DA, DB, DO><<<static_cast<int>((rem + 255) / 256), 256, 0, s>>>(a, b, out, done, n, sign);
    MC_CHECK_LAUNCH("axpb_scalar_kernel launch");
  }
  return MC_OK;
}

static int32_t dispatch_axpb(const void* a, int da, const void* b, int db, void* out, int dout, int64_t n, float sign,
                             cudaStream_t s, const char* who) {
  MC_CHECK_ARG(n >= 0, "%s: negative element count", who);
  if (n == 0) return MC_OK;  // empty tensors have null data pointers
  MC_CHECK_ARG(a && b && out, "%s: null pointer", who);
#define MC_CASE(A, B, O) \
  if (da == A && db == B && dout == O) return launch_axpb<A, B, O>(a, b, out, n, sign, s);
  MC_CASE(MC_BF16, MC_F32, MC_F32)    // Wan hit: bf16 patch-embed output + fp32 residual
  MC_CASE(MC_F32, MC_BF16, MC_F32)    // Wan miss: fp32 stream - bf16 ori_x
  MC_CASE(MC_F32, MC_F32, MC_F32)
  MC_CASE(MC_BF16, MC_BF16, MC_BF16)  // FLUX / Hunyuan (all bf16)
  MC_CASE(MC_BF16, MC_BF16, MC_F32)
  MC_CASE(MC_F32, MC_BF16, MC_BF16)
  MC_CASE(MC_BF16, MC_F32, MC_BF16)
  MC_CASE(MC_F32, MC_F32, MC_BF16)
#undef MC_CASE
  set_error("%s: unsupported dtype combination (%d, %d) -> %d", who, da, db, dout);
  return MC_ERR_INVALID;
}

// ---- CFG combine of the caller loop: out = uncond + g * (cond - uncond), each op rounded separately like torch eager --------
__global__ void __launch_bounds__(256) cfg_combine_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, float g,
                                                          float* __restrict__ out, int64_t n_groups) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_groups; i += stride) {
    float c[8], u[8];
    Elem<MC_F32>::load8(cond, i * 8, c);
    Elem<MC_F32>::load8(uncond, i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = __fadd_rn(u[j], __fmul_rn(g, __fsub_rn(c[j], u[j])));
    Elem<MC_F32>::store8(out, i * 8, c);
  }
}
__global__ void cfg_combine_tail_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, float g,
                                        float* __restrict__ out, int64_t begin, int64_t n) {
  const int64_t i = begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(uncond[i], __fmul_rn(g, __fsub_rn(cond[i], uncond[i])));
}

// ---- K3: per-row norms / cosine, single pass ---------------------------------------------------------------
// One warp per row; lanes stride over 8-element groups. Optionally also forms cur = xo - xi on the fly and stores it.
template <int DCUR, int DPREV, bool FUSE_SUB>
__global__ void __launch_bounds__(256) stats_kernel(const void* __restrict__ cur_or_xo, const void* __restrict__ xi_bf16,
                                                    void* __restrict__ r_out, const void* __restrict__ prev, int64_t rows, int cols,
                                                    double denom_eps, double* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
  const int groups = cols >> 3;
  double acc_ratio = 0.0, acc_ratio2 = 0.0, acc_cos = 0.0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * warps_per_block + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t base = row * cols;
    float cc = 0.f, pp = 0.f, cp = 0.f;
    // kBatch groups per lane are loaded before any of them is reduced: 2 x kBatch independent 128/256-bit loads in flight
    constexpr int kBatch = 3;
    for (int g0 = lane; g0 < groups; g0 += 32 * kBatch) {
      float c[kBatch][8], p[kBatch][8];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int g = g0 + 32 * u;
        if (g < groups) {
          Elem<DCUR>::load8(cur_or_xo, base + g * 8, c[u]);
          if (FUSE_SUB) {
            float xi[8];
            Elem<MC_BF16>::load8(xi_bf16, base + g * 8, xi);
#pragma unroll
            for (int j = 0; j < 8; ++j) c[u][j] = c[u][j] - xi[j];
          }
          Elem<DPREV>::load8(prev, base + g * 8, p[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int g = g0 + 32 * u;
        if (g < groups) {
          if (FUSE_SUB) Elem<MC_F32>::store8(r_out, base + g * 8, c[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            cc = fmaf(c[u][j], c[u][j], cc);
            pp = fmaf(p[u][j], p[u][j], pp);
            cp = fmaf(c[u][j], p[u][j], cp);
          }
        }
      }
    }
    cc = warp_sum(cc);
    pp = warp_sum(pp);
    cp = warp_sum(cp);
    if (lane == 0) {
      const float n_cur = sqrtf(cc), n_prev = sqrtf(pp);
      const float ratio = n_cur / (n_prev + static_cast<float>(denom_eps));
      // F.cosine_similarity(eps=1e-8): sum((a/max(|a|,eps)) * (b/max(|b|,eps)))
      const float cosv = cp / (fmaxf(n_cur, 1e-8f) * fmaxf(n_prev, 1e-8f));
      acc_ratio += static_cast<double>(ratio);
      acc_ratio2 += static_cast<double>(ratio) * static_cast<double>(ratio);
      acc_cos += static_cast<double>(1.0f - cosv);
    }
  }
  __shared__ double sh[3][8];
  if (lane == 0) {
    sh[0][warp] = acc_ratio;
    sh[1][warp] = acc_ratio2;
    sh[2][warp] = acc_cos;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < warps_per_block; ++w) t += sh[threadIdx.x][w];
    atomicAdd(&stats[threadIdx.x], t);
  }
}

__global__ void stats_init_kernel(double* stats, double rows) {
  if (threadIdx.x < 3) stats[threadIdx.x] = 0.0;
  if (threadIdx.x == 3) stats[3] = rows;
}

template <int DCUR, int DPREV, bool FUSE>
static int32_t launch_stats(const void* cur, const void* xi, void* r_out, const void* prev, int64_t rows, int cols, double eps,
                            double* stats, cudaStream_t s) {
  stats_init_kernel<<<1, 32, 0, s>>>(stats, static_cast<double>(rows));
  MC_CHECK_LAUNCH("stats_init_kernel launch");
  const int threads = 256, wpb = threads / 32;
  int64_t want = (rows + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const int grid = static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
  stats_kernel<DCUR, DPREV, FUSE><<<grid, threads, 0, s>>>(cur, xi, r_out, prev, rows, cols, eps, stats);
  MC_CHECK_LAUNCH("stats_kernel launch");
  return MC_OK;
}

}  // namespace mc

extern "C" {

int32_t mc_cache_hit_add(const void* x, int32_t x_dtype, const void* r, int32_t r_dtype, void* out, int32_t out_dtype, int64_t n,
                         void* stream) {
  return mc::dispatch_axpb(x, x_dtype, r, r_dtype, out, out_dtype, n, 1.0f, static_cast<cudaStream_t>(stream), "mc_cache_hit_add");
}

int32_t mc_residual_sub(const void* x_out, int32_t xo_dtype, const void* x_in, int32_t xi_dtype, void* r, int32_t r_dtype, int64_t n,
                        void* stream) {
  return mc::dispatch_axpb(x_out, xo_dtype, x_in, xi_dtype, r, r_dtype, n, -1.0f, static_cast<cudaStream_t>(stream),
                           "mc_residual_sub");
}

int32_t mc_cfg_combine(const float* cond, const float* uncond, float guide_scale, float* out, int64_t n, void* stream) {
  MC_CHECK_ARG(n >= 0, "mc_cfg_combine: negative element count");
  if (n == 0) return MC_OK;
  MC_CHECK_ARG(cond && uncond && out, "mc_cfg_combine: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
  const int64_t groups = (al32(cond) && al32(uncond) && al32(out)) ? n / 8 : 0;
  if (groups > 0) {
    int64_t want = (groups + 255) / 256;
    const int64_t cap = static_cast<int64_t>(mc::num_sms()) * 8;
    mc::cfg_combine_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(cond, uncond, guide_scale, out, groups);
    MC_CHECK_LAUNCH("cfg_combine_kernel launch");
  }
  if (groups * 8 < n) {
    const int64_t rem = n - groups * 8;
    mc::cfg_combine_tail_kernel<<<static_cast<int>((rem + 255) / 256), 256, 0, s>>>(cond, uncond, guide_scale, out, groups * 8, n);
    MC_CHECK_LAUNCH("cfg_combine_tail_kernel launch");
  }
  return MC_OK;
}

int32_t mc_residual_stats(const void* r_cur, int32_t cur_dtype, const void* r_prev, int32_t prev_dtype, int64_t rows, int32_t cols,
                          double denom_eps, double* stats_dev, void* stream) {
  MC_CHECK_ARG(r_cur && r_prev && stats_dev, "mc_residual_stats: null pointer");
  MC_CHECK_ARG(rows >= 1 && cols >= 8 && cols % 8 == 0, "mc_residual_stats: rows=%lld cols=%d (cols must be a multiple of 8)",
               static_cast<long long>(rows), cols);
  MC_CHECK_ARG((reinterpret_cast<uintptr_t>(r_cur) & 31u) == 0 && (reinterpret_cast<uintptr_t>(r_prev) & 31u) == 0,
               "mc_residual_stats: pointers must be 32-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cur_dtype == MC_F32 && prev_dtype == MC_F32)
    return mc::launch_stats<MC_F32, MC_F32, false>(r_cur, nullptr, nullptr, r_prev, rows, cols, denom_eps, stats_dev, s);
  if (cur_dtype == MC_BF16 && prev_dtype == MC_BF16)
    return mc::launch_stats<MC_BF16, MC_BF16, false>(r_cur, nullptr, nullptr, r_prev, rows, cols, denom_eps, stats_dev, s);
  mc::set_error("mc_residual_stats: unsupported dtypes (%d, %d)", cur_dtype, prev_dtype);
  return MC_ERR_INVALID;
}

int32_t mc_residual_sub_stats(const void* x_out, int32_t xo_dtype, const void* x_in, int32_t xi_dtype, void* r, const void* r_prev,
                              int64_t rows, int32_t cols, double denom_eps, double* stats_dev, void* stream) {
  MC_CHECK_ARG(x_out && x_in && r && r_prev && stats_dev, "mc_residual_sub_stats: null pointer");
  MC_CHECK_ARG(xo_dtype == MC_F32 && xi_dtype == MC_BF16, "mc_residual_sub_stats: only fp32 - bf16 -> fp32 (the Wan stream) is built");
  MC_CHECK_ARG(rows >= 1 && cols >= 8 && cols % 8 == 0, "mc_residual_sub_stats: cols must be a multiple of 8");
  MC_CHECK_ARG((reinterpret_cast<uintptr_t>(x_out) & 31u) == 0 && mc::aligned16(x_in) && (reinterpret_cast<uintptr_t>(r) & 31u) == 0 &&
                   (reinterpret_cast<uintptr_t>(r_prev) & 31u) == 0,
               "mc_residual_sub_stats: fp32 pointers must be 32-byte aligned, bf16 16-byte");
  return mc::launch_stats<MC_F32, MC_F32, true>(x_out, x_in, r, r_prev, rows, cols, denom_eps, stats_dev,
                                                static_cast<cudaStream_t>(stream));
}

}  // extern "C"
