// Thin inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences) and the UMMA descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables (the same ones
// cute/arch/mma_sm100_desc.hpp encodes); nothing here depends on CUTLASS.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
#ifndef MC_MBAR_TIMEOUT_CYCLES
#define MC_MBAR_TIMEOUT_CYCLES (4000000000ll)  // ~2 s at 1.9 GHz
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > MC_MBAR_TIMEOUT_CYCLES) {
      printf("mbar_wait timeout: block (%d,%d,%d) thread %d bar smem+%u parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// One lane of a fully converged warp. tcgen05.mma / commit and the TMA instructions take their operands from UNIFORM registers: under
// a divergent `if (lane == 0)` ptxas has to wrap every one of them in a uniformisation loop (ELECT / PLOP3 / BRA.U.ANY, ~10
// instructions and ~100 cycles per MMA — enough to make the single issuing thread the bottleneck of a kernel whose MMAs take 64
// cycles each). Issued under elect.sync from warp-uniform control flow they are single instructions. The elected lane is the same on
// every call (the lowest active one), which tcgen05.commit relies on.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// proxies / fences
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic-proxy smem writes -> visible to async proxy (UMMA/TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load: c0 = element index along the contiguous dimension, c1 = row index.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMEM allocation (one warp, .sync.aligned)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile stored as [rows][64 bf16] (128-byte rows)
// with the 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B). 8-row groups are 1024 B apart (SBO).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (sm_100)      bits [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major operand tile (the "N x K" B operand stored with N contiguous, e.g. V [kv, d] row-major for O += P V): TMA boxes of
// [K rows][64 bf16] (128-byte rows, 128-byte swizzle). In 16-byte units the canonical layout is ((8,n),(8,k)):((1,LBO),(8,SBO)):
// 64 MN-elements contiguous, the next 64 MN-elements LBO bytes further (= one TMA box), 8 K-rows 128 B apart, the next 8 K-rows
// SBO = 1024 B further. One UMMA K-step (16 K-rows) advances the start address by 2048 B.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, dense, no negate.
//   [4,6) D fmt (1=f32)  [7,10) A fmt (1=bf16)  [10,13) B fmt (1=bf16)  [15] A major (0=K)  [16] B major (0=K)
//   [17,23) N>>3         [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// same with the B operand MN-major (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_bmn(int M, int N) { return umma_idesc_bf16_f32(M, N) | (1u << 16); }

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A (bf16, K-major) lives in TMEM — lane = row, each 32-bit column holds two consecutive
// K elements (16 bf16 of one UMMA_K step = 8 columns).
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all tcgen05 async ops issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMEM <-> registers. 32x32b: lane i of the warp reads TMEM lane (base_lane + i), N consecutive 32-bit columns.
// taddr = (lane << 16) | column.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Register reallocation between the warpgroups of a CTA (all 4 warps of a warpgroup execute it): the data-path warpgroups give
// registers up, the compute warpgroups take them. N a multiple of 8 in [24, 256].
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ------------------------------------------------------------------------------------------------
// misc math / memory
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Blackwell packed-fp32 and 3-input ALU ops (halve the issue slots of the softmax inner loop)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// 2^a, 2^b on the FMA / ALU pipes instead of the MUFU (FlashAttention-4's trick): Cody-Waite split x = floor(x) + f with the
// magic-constant add rounded toward -inf, a cubic for 2^f on [0, 1) (max relative error 8.8e-5 — 22x below the bf16 rounding P
// gets anyway), and the integer part added into the exponent field. 10 SASS instructions per PAIR (2 FMNMX, FADD2.RM, FADD2,
// 4 FFMA2, 2 LEA) against 2 MUFU.EX2 that occupy the 16-lane special-function unit for 8 cycles per warp each.
// Inputs below -126 (masked columns: -inf) give 2^-126 * [1, 2) instead of 0: finite, and below anything the row sum can see.
__device__ __forceinline__ void ex2_emul_pair(float& a, float& b) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23
  a = fmaxf(a, -126.0f);
  b = fmaxf(b, -126.0f);
  uint64_t x2, t2, fi2, f2, p2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(a), "f"(b));
  uint64_t magic2, nmagic2, none2, c3, c2, c1, c0;
  asm("mov.b64 %0, {%1, %1};" : "=l"(magic2) : "f"(kMagic));
  asm("mov.b64 %0, {%1, %1};" : "=l"(nmagic2) : "f"(-kMagic));
  asm("mov.b64 %0, {%1, %1};" : "=l"(none2) : "f"(-1.0f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c3) : "f"(0.077119089663028717041015625f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(0.227564394474029541015625f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c1) : "f"(0.695146143436431884765625f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c0) : "f"(1.0f));
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(x2), "l"(magic2));    // low mantissa bits of t = floor(x)
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(fi2) : "l"(t2), "l"(nmagic2));       // floor(x) as a float
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2) : "l"(fi2), "l"(none2), "l"(x2));  // f = x - floor(x) in [0, 1)
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(c3), "l"(f2), "l"(c2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(f2), "l"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(f2), "l"(c0));
  float ta, tb, pa, pb;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(ta), "=f"(tb) : "l"(t2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(pa), "=f"(pb) : "l"(p2));
  a = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
  b = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {  // streaming 128-bit load, no L1 allocation
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {  // streaming 128-bit store
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 256-bit global accesses (sm_100+): one instruction per lane moves 8 fp32, so a warp touches 1 KB contiguously and every
// 32-byte sector is requested exactly once (two .v4 accesses per lane would request each sector twice from L2).
__device__ __forceinline__ void ld_nc_v8_f32(const float* p, float (&f)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(p));
}
__device__ __forceinline__ void ld_v8_f32(const float* p, float (&f)[8]) {  // coherent form: the buffer may be written by this kernel
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void st_na_v8_f32(float* p, const float (&f)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]),
               "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7])
               : "memory");
}
}  // namespace ptx
