/* The boundary is a true C ABI: this file is compiled with gcc -std=c99 against include/b200sense.h,
 * links libb2sense.so and exercises the argument validation paths (no GPU needed). */
#include <stdio.h>
#include <string.h>
#include "b200sense.h"

int main(void) {
  int fails = 0;
  if (b2s_version() < 100) { printf("bad version\n"); fails++; }
  if (b2s_has_fused_plan(200, 200) != 1 || b2s_has_fused_plan(256, 256) != 1 || b2s_has_fused_plan(64, 64) != 0) { printf("plan query\n"); fails++; }
  if (b2s_scratch_bytes(1, 2, 3, 200, 200) != 0 || b2s_scratch_bytes(1, 2, 3, 10, 6) != (size_t)2 * 3 * 10 * 6 * 8) { printf("scratch\n"); fails++; }
  if (b2s_fft2c(NULL, NULL, 1, 200, 200, 0, 9, NULL) != B2S_EINVAL) { printf("norm check\n"); fails++; }
  if (strstr(b2s_last_error(), "bad argument") == NULL) { printf("error text: %s\n", b2s_last_error()); fails++; }
  if (b2s_fft2c(NULL, NULL, 1, 200, 200, 0, B2S_NORM_ORTHO, NULL) != B2S_EINVAL) { printf("null check\n"); fails++; }
  if (b2s_fft2c(NULL, NULL, 0, 200, 200, 0, B2S_NORM_ORTHO, NULL) != B2S_OK) { printf("empty batch\n"); fails++; }
  if (b2s_sens_expand(NULL, NULL, NULL, NULL, NULL, NULL, B2S_EXPAND_DC, 0, 15, 10, 200, 200, B2S_NORM_ORTHO, NULL, 0, NULL) != B2S_OK) { printf("empty expand\n"); fails++; }
  if (b2s_sens_expand((const float*)8, (const float*)8, (float*)8, NULL, NULL, NULL, B2S_EXPAND_DC, 1, 1, 1, 200, 200, B2S_NORM_ORTHO, NULL, 0, NULL) != B2S_EINVAL) { printf("dc needs ref\n"); fails++; }
  if (b2s_sens_reduce((const float*)8, (const float*)8, (float*)8, NULL, NULL, B2S_REDUCE_MASK, 0, 1, 1, 1, 200, 200, B2S_NORM_ORTHO, NULL, 0, NULL) != B2S_EINVAL) { printf("mask needed\n"); fails++; }
  if (b2s_normal_op((const float*)8, (const float*)8, (const uint8_t*)8, (const float*)8, (float*)8, 1, 1, 1, 128, 128, NULL) != B2S_EUNSUPPORTED) { printf("normal op shape\n"); fails++; }
  if (b2s_launch_count(1) != 0ULL) { printf("no kernel may have been launched\n"); fails++; }
  printf(fails ? "FAIL %d\n" : "OK\n", fails);
  return fails;
}
