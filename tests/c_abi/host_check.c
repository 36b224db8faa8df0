/* A plain-C host of libhwgat_b200: proves that include/hwgat_b200.h is a C header (no C++ types across the boundary)
 * and that the housekeeping entry points work without a GPU.  Built and run by tests/test_host.py. */
#include <stdio.h>
#include <string.h>

#include "hwgat_b200.h"

int main(void) {
  int ok = 1;
  printf("version %d\n", hwgat_version());
  ok &= hwgat_version() >= 11;
  ok &= strstr(hwgat_error_string(HWGAT_ERR_UNSUPPORTED), "no fallback") != NULL;
  ok &= strcmp(hwgat_error_string(HWGAT_OK), "ok") == 0;
  /* size queries are pure host arithmetic */
  ok &= hwgat_attn_workspace_bytes(2, 4, 64, 128, 2, HWGAT_BF16, 0) == (size_t)(384 * 128 + 2 * 192 * 16) * 2;
  ok &= hwgat_attn2_workspace_bytes(2, 4, 64, 128, 2, 0, 1) > 0;
  ok &= hwgat_ln_pool_scratch_bytes(512, 1024, 512) > 0;
  ok &= hwgat_ln_pool_scratch_bytes(4096, 1024, 512) == 0;
  /* argument errors are reported before any device work */
  ok &= hwgat_attn_fwd(NULL, NULL, NULL, NULL, -1.0f, NULL, NULL, 0, 1, 4, 64, 128, 2, 8, 2, 0, HWGAT_LAYOUT_BFKD, HWGAT_F32,
                       NULL) == HWGAT_ERR_UNSUPPORTED;
  ok &= hwgat_attn2_fwd(NULL, NULL, NULL, NULL, -1.0f, NULL, NULL, NULL, 0, 1, 4, 64, 128, 2, 48, 2, 0, HWGAT_LAYOUT_BFKD,
                        0.0f, 0ull, 0ull, NULL) == HWGAT_ERR_UNSUPPORTED;
  ok &= hwgat_attn2_fwd(NULL, NULL, NULL, NULL, -1.0f, NULL, NULL, NULL, 0, 1, 4, 64, 128, 2, 32, 2, 0, HWGAT_LAYOUT_BFKD,
                        0.0f, 0ull, 0ull, NULL) == HWGAT_ERR_NULL;
  ok &= hwgat_smooth_ce_fwd(NULL, NULL, NULL, NULL, NULL, 4, 10, 0.01f, NULL) == HWGAT_ERR_NULL;
  printf(ok ? "c abi ok\n" : "c abi FAILED\n");
  return ok ? 0 : 1;
}
