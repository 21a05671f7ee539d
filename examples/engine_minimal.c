/* Minimal C caller of the whole-path engine (include/udb.h): what a non-Python host links against.
 *   gcc -I include examples/engine_minimal.c -L unidepth_b200 -ludb -Wl,-rpath,$PWD/unidepth_b200 -o engine_minimal
 * Without packed weights it only exercises the host-side part of the ABI: handle life cycle, the
 * geometry of `UniDepthV2.infer` (unidepthv2.py:36-77, 247-262) and the error reporting.  A real caller
 * registers the packed tensors (udb_set_weight), sizes the workspace and calls udb_infer_v2 per batch
 * (see INTEGRATION.md section B). */
#include <stdio.h>
#include <string.h>

#include "udb.h"

int main(void) {
  udb_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.embed_dim = 1024; cfg.depth = 24; cfg.enc_heads = 16; cfg.pos_grid = 37;
  cfg.taps[0] = 6; cfg.taps[1] = 12; cfg.taps[2] = 18; cfg.taps[3] = 24;
  cfg.hidden = 512; cfg.dec_heads = 8; cfg.expansion = 4; cfg.out_dim = 64; cfg.n_stages = 3;
  cfg.dec_depths[0] = cfg.dec_depths[1] = cfg.dec_depths[2] = 2;
  cfg.ratio_min = 0.5; cfg.ratio_max = 2.5; cfg.pixels_min = 200000; cfg.pixels_max = 600000;

  udb_engine* eng = NULL;
  if (udb_create(&cfg, &eng)) { fprintf(stderr, "udb_create: %s\n", udb_last_error()); return 1; }

  const int shapes[4][2] = {{480, 640}, {1024, 1536}, {480, 1600}, {1000, 400}};
  for (int i = 0; i < 4; ++i) {
    udb_geometry_t g;
    if (udb_geometry(eng, shapes[i][0], shapes[i][1], -1, &g)) { fprintf(stderr, "%s\n", udb_last_error()); return 1; }
    printf("%dx%d -> pad l%d r%d t%d b%d, network %dx%d (grid %dx%d), factor %.6f\n", shapes[i][0], shapes[i][1], g.pad_l,
           g.pad_r, g.pad_t, g.pad_b, g.net_h, g.net_w, g.gh, g.gw, g.factor);
  }
  /* nothing registered yet: the engine refuses instead of computing */
  if (udb_workspace_bytes(eng, 8, 480, 640, -1) == 0) printf("as expected: %s\n", udb_last_error());
  udb_destroy(eng);
  printf("udb version %d\n", udb_version());
  return 0;
}
