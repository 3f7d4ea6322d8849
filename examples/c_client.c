/* Minimal plain-C client of the C ABI (include/bsvd_b200.h): what a non-Python host would link.
 *
 *   gcc -std=c99 -Iinclude examples/c_client.c -Lbsvd_b200/lib -lbsvd_b200 -Wl,-rpath,$PWD/bsvd_b200/lib -o c_client
 *
 * With a B200 it creates a BSVD-64 handle and reports the expected weight shapes; without a GPU
 * bsvd_create fails loudly ("no CPU fallback") and the program says so.  It never computes on the
 * host.  Device buffers (cudaMalloc) and the weight upload are the caller's business:
 *   bsvd_set_weights(h, layer, w_oihw_host, bias_host, out_ch, in_ch)   for layer = 0..31
 *   bsvd_forward_clip(h, d_in, d_noise_map_or_NULL, d_out, T, in_c, H, W, stream)
 */
#include <stdio.h>
#include <string.h>

#include "bsvd_b200.h"

int main(void) {
  bsvd_config cfg;
  bsvd_handle* h = NULL;
  int layer;
  memset(&cfg, 0, sizeof(cfg));
  cfg.chns[0] = 64; cfg.chns[1] = 128; cfg.chns[2] = 256;   /* options/test/bsvd_c64.yml:85-93 */
  cfg.mid_ch = 64; cfg.interm_ch = 64; cfg.in_ch = 4; cfg.out_ch = 3;
  cfg.act_relu6 = 1; cfg.norm_none = 1; cfg.precision = BSVD_PREC_FP16; cfg.device = -1;
  printf("%s\n", bsvd_version());
  if (bsvd_create(&cfg, &h) != 0) {
    printf("bsvd_create failed: %s\n", bsvd_last_error());
    return 0;   /* expected on a machine without a B200 */
  }
  for (layer = 0; layer < BSVD_NUM_LAYERS; ++layer) {
    int co = 0, ci = 0, st = 0;
    if (bsvd_layer_shape(h, layer, &co, &ci, &st) != 0) {
      printf("bsvd_layer_shape failed: %s\n", bsvd_last_error());
      return 1;
    }
    printf("layer %2d: weight [%d,%d,3,3] stride %d\n", layer, co, ci, st);
  }
  bsvd_destroy(h);
  return 0;
}
