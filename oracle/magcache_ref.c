/* ORACLE — test infrastructure only. Plain-C restatement of the MagCache host logic and of the three HBM-bound cache
 * operations, used (a) by tests/ as a second, independent checker of libmagcache_b200 and (b) by bench.py as the CPU baseline
 * ("port") for the cache kernels. Build: `make -C oracle` -> oracle/_build/libmagcache_ref.so. Never linked by the product.
 *
 *   ref_nearest_interp      MagCache4Wan2.1/magcache_generate.py:27-34
 *   ref_ctrl_mask           MagCache4Wan2.1/magcache_generate.py:277-292,306-311 ; MagCache4FLUX/magcache_flux.py:326-338 ;
 *                           MagCache4HunyuanVideo/magcache_sample_video.py:88-102
 *   ref_hit_add_bf16_f32    `x = x + residual_x`      magcache_generate.py:295  (bf16 + fp32 -> fp32)
 *   ref_sub_f32_bf16        `residual_x = x - ori_x`  magcache_generate.py:299  (fp32 - bf16 -> fp32)
 *   ref_stats_f32           calibration statistics    magcache_generate.py:167-169
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int ref_nearest_interp(const double* src, int L, double* dst, int T) {
  if (T == 1) {
    dst[0] = src[L - 1];
    return 0;
  }
  double scale = (double)(L - 1) / (double)(T - 1);
  for (int i = 0; i < T; ++i) dst[i] = src[(long)nearbyint((double)i * scale)];
  return 0;
}

/* family: 0 wan2.1 (per-branch, <, int(n*R)) ; 1 flux (scalar, <=, int(R*n+.5), veto 11 of 28) ; 2 hunyuan (scalar, <=, int(R*n)) */
int ref_ctrl_mask(int family, const double* ratios, int num_steps, double thresh, int K, double R, int calls, uint8_t* mask) {
  double acc_ratio[2] = {1.0, 1.0}, acc_err[2] = {0.0, 0.0};
  int acc_steps[2] = {0, 0};
  int cnt = 0;
  int start = (family == 1) ? (int)(R * (double)num_steps + 0.5) : (int)((double)num_steps * R);
  for (int c = 0; c < calls; ++c) {
    int skip = 0;
    if (cnt >= start) {
      int i = (family == 0) ? (cnt % 2) : 0;
      acc_ratio[i] = acc_ratio[i] * ratios[cnt];
      acc_steps[i] += 1;
      acc_err[i] += fabs(1.0 - acc_ratio[i]);
      int ok = (family == 0) ? (acc_err[i] < thresh) : (acc_err[i] <= thresh);
      ok = ok && (acc_steps[i] <= K);
      if (family == 1) ok = ok && ((long)nearbyint((double)cnt * (27.0 / (double)(num_steps - 1))) != 11);
      if (ok) {
        skip = 1;
      } else {
        acc_err[i] = 0.0;
        acc_steps[i] = 0;
        acc_ratio[i] = 1.0;
      }
    }
    mask[c] = (uint8_t)skip;
    cnt += 1;
    if (cnt >= num_steps) {
      cnt = 0;
      acc_ratio[0] = acc_ratio[1] = 1.0;
      acc_err[0] = acc_err[1] = 0.0;
      acc_steps[0] = acc_steps[1] = 0;
    }
  }
  return 0;
}

void ref_hit_add_bf16_f32(const uint16_t* x, const float* r, float* out, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = bf16_to_f32(x[i]) + r[i];
}

void ref_sub_f32_bf16(const float* xo, const uint16_t* xi, float* out, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = xo[i] - bf16_to_f32(xi[i]);
}

/* out[0..2] = mean ratio, unbiased std of ratio, mean (1 - cos) ; fp32 per-row arithmetic like torch, fp64 across rows */
void ref_stats_f32(const float* cur, const float* prev, int64_t rows, int cols, double denom_eps, double* out) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, s2)
  for (int64_t r = 0; r < rows; ++r) {
    const float* c = cur + r * cols;
    const float* p = prev + r * cols;
    double cc = 0.0, pp = 0.0, cp = 0.0;
    for (int j = 0; j < cols; ++j) {
      cc += (double)c[j] * c[j];
      pp += (double)p[j] * p[j];
      cp += (double)c[j] * p[j];
    }
    float nc = (float)sqrt(cc), np_ = (float)sqrt(pp);
    float ratio = nc / (np_ + (float)denom_eps);
    float cosv = (float)cp / (fmaxf(nc, 1e-8f) * fmaxf(np_, 1e-8f));
    s0 += ratio;
    s1 += (double)ratio * ratio;
    s2 += 1.0 - cosv;
  }
  double n = (double)rows;
  out[0] = s0 / n;
  out[1] = rows > 1 ? sqrt(fmax((s1 - s0 * s0 / n) / (n - 1.0), 0.0)) : NAN;
  out[2] = s2 / n;
}
