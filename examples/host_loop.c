/* A host that is not Python: the denoising loop of one video on libmagcache_b200.so through the C ABI alone — the controller
 * (MagCache4Wan2.1/magcache_generate.py:277-292, :306-311) as mc_ctrl_*, every patched forward (:229-275, :293-305) as ONE
 * mc_dit_forward call. This file is plain C (the header carries no C++): `gcc -std=c99 -Iinclude examples/host_loop.c
 * -Lmagcache_b200 -lmagcache_b200`; tests/test_c_host_cpu.py compiles it and runs its host-only part (`--plan`), which needs no GPU:
 * the skip schedule of the run and the launch plan of a miss / a hit forward. With device buffers supplied by the embedding
 * application (`run_video`) the same code drives the GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "magcache_b200.h"

/* One video: `steps` denoising steps = 2 * steps patched forwards (cond, uncond, cond, ...; wan_magcache.py:296-299). All pointers are
 * device pointers owned by the caller; residual[0 / 1] are the reference's residual_cache[cnt % 2]. Returns the forwards skipped. */
static int run_video(mc_dit* dit, mc_ctrl* ctrl, int steps, const float* latent, const double* const* t_dev, const void* ctx, const void* ctx_null,
                     float* const residual[2], float* const noise_pred[2], void* stream) {
  int skipped = 0;
  for (int call = 0; call < 2 * steps; ++call) {
    int32_t skip = 0;
    if (mc_ctrl_step(ctrl, &skip, NULL) != MC_OK) {
      fprintf(stderr, "controller: %s\n", mc_last_error());
      return -1;
    }
    if (mc_dit_forward(dit, latent, t_dev[call / 2], (call % 2) ? ctx_null : ctx, skip, residual[call % 2], noise_pred[call % 2], stream) != MC_OK) {
      fprintf(stderr, "forward %d: %s\n", call, mc_last_error());
      return -1;
    }
    skipped += skip;
    /* the caller's CFG combine + scheduler update go here: mc_cfg_step(noise_pred[0], noise_pred[1], guide_scale, latent, ...) */
  }
  return skipped;
}

int main(int argc, char** argv) {
  if (argc < 2 || strcmp(argv[1], "--plan") != 0) {
    (void)run_video;
    fprintf(stderr, "usage: %s --plan   (host-only: skip schedule + launch plan; the device loop is run_video())\n", argv[0]);
    return 2;
  }
  printf("abi %d\n", (int)mc_abi_version());

  /* ---- the controller of a 10-step run with a constant magnitude ratio: E012 K2 R0.2 (magcache_generate.py:896-919) */
  enum { STEPS = 10 };
  double table[2 * STEPS];
  for (int i = 0; i < 2 * STEPS; ++i) table[i] = i < 2 ? 1.0 : 0.97;
  mc_ctrl_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_steps = 2 * STEPS, cfg.branches = 2, cfg.K = 2, cfg.cmp = MC_CMP_LT, cfg.retention_mode = MC_RETAIN_FLOOR, cfg.veto_index = -1;
  cfg.thresh = 0.12, cfg.retention_ratio = 0.2, cfg.mag_ratios = table;
  mc_ctrl* ctrl = mc_ctrl_create(&cfg, 0);
  if (!ctrl) {
    fprintf(stderr, "mc_ctrl_create: %s\n", mc_last_error());
    return 1;
  }
  printf("skip ");
  for (int call = 0; call < 2 * STEPS; ++call) {
    int32_t skip = 0;
    if (mc_ctrl_step(ctrl, &skip, NULL) != MC_OK) return 1;
    printf("%d", (int)skip);
  }
  printf("\n");
  mc_ctrl_destroy(ctrl);

  /* ---- the forward's launch plan (no weights are read to plan: stand-in addresses) */
  enum { LAYERS = 2 };
  mc_dit_dims dims = {256, 512, 2, LAYERS, 16, 16, 256, 128, 32, 1e-6f};
  static mc_dit_block blocks[LAYERS];
  mc_dit_weights w;
  uintptr_t next = (uintptr_t)1 << 44;
  const void** p = (const void**)blocks;
  for (size_t i = 0; i < (sizeof blocks) / (sizeof(void*)); ++i, next += (uintptr_t)1 << 32) p[i] = (const void*)next;
  p = (const void**)&w;
  for (size_t i = 0; i < (sizeof w) / (sizeof(void*)); ++i, next += (uintptr_t)1 << 32) p[i] = (const void*)next;
  w.blocks = blocks;
  mc_dit* dit = mc_dit_create(&dims, &w);
  if (!dit) {
    fprintf(stderr, "mc_dit_create: %s\n", mc_last_error());
    return 1;
  }
  int64_t bytes = 0;
  if (mc_dit_workspace_bytes(dit, 2, 4, 6, &bytes) != MC_OK) return 1;
  printf("workspace %lld\n", (long long)bytes);
  if (mc_dit_bind(dit, 2, 4, 6, (void*)((uintptr_t)1 << 40), bytes, (const float*)next) != MC_OK) {
    fprintf(stderr, "mc_dit_bind: %s\n", mc_last_error());
    return 1;
  }
  for (int skip = 0; skip <= 1; ++skip) {
    int64_t need = 0;
    if (mc_dit_plan(dit, skip, NULL, 0, &need) != MC_OK) return 1;
    char* text = (char*)malloc((size_t)need);
    if (mc_dit_plan(dit, skip, text, need, &need) != MC_OK) return 1;
    int lines = 0;
    for (const char* c = text; *c; ++c) lines += *c == '\n';
    printf("plan skip=%d lines=%d first=%.*s\n", skip, lines, (int)strcspn(text, "\n"), text);
    free(text);
  }
  mc_dit_destroy(dit);
  return 0;
}
