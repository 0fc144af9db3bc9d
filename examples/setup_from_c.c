/* A plain-C consumer of libfemb200's host-side setup: builds the multigrid hierarchy of an n x n x n voxel cube through
 * femb_setup_hierarchy_host (what Object.__init__ of /root/reference/exp1/quasi_simulation/object.py:116-697 computes) and
 * prints one line per level.  No CUDA device is needed for this call; the arrays it returns are what femb_*'s device entry
 * points take after a cudaMemcpy.
 *
 *   gcc -std=c99 -I include examples/setup_from_c.c -L fem-simulation_b200 -lfemb200 -Wl,-rpath,$PWD/fem-simulation_b200 -o setup_from_c
 *   ./setup_from_c 12 3          # cube edge in voxels, number of levels
 */
#include <stdio.h>
#include <stdlib.h>

#include "femb200.h"

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8;
  const int levels = argc > 2 ? atoi(argv[2]) : 2;
  const float dx = 1.0f / (float)n;
  const int p = n + 1, n_vert = p * p * p, n_hex = n * n * n;
  float* x = (float*)malloc(sizeof(float) * 3 * (size_t)n_vert);
  int* hex = (int*)malloc(sizeof(int) * 8 * (size_t)n_hex);
  /* the reference's numbering (object.py:28-45): lattice points x-fastest, local node a = 4 bx + 2 by + bz */
  for (int iz = 0; iz < p; ++iz)
    for (int iy = 0; iy < p; ++iy)
      for (int ix = 0; ix < p; ++ix) {
        float* q = x + 3 * (size_t)(ix + p * (iy + p * iz));
        q[0] = (float)ix * dx; q[1] = (float)iy * dx; q[2] = (float)iz * dx;
      }
  for (int iz = 0; iz < n; ++iz)
    for (int iy = 0; iy < n; ++iy)
      for (int ix = 0; ix < n; ++ix)
        for (int a = 0; a < 8; ++a)
          hex[8 * (size_t)(ix + n * (iy + n * iz)) + a] = (ix + ((a >> 2) & 1)) + p * ((iy + ((a >> 1) & 1)) + p * (iz + (a & 1)));
  FembHierarchy* h = NULL;
  const int rc = femb_setup_hierarchy_host(n_vert, n_hex, x, hex, dx, levels, 0, NULL, &h);
  if (rc != FEMB_OK) {
    fprintf(stderr, "femb_setup_hierarchy_host: %d (%s)\n", rc, femb_last_error());
    return 1;
  }
  printf("abi %d levels %d\n", femb_abi_version(), h->n_levels);
  for (int l = 0; l < h->n_levels; ++l) {
    const FembHierarchyLevel* L = &h->level[l];
    double wsum = 0.0;
    for (int i = 0; i < L->n_entries; ++i) wsum += L->us_hat_w[i];
    printf("level %d: %d vertices, %d hexahedra, %d blocks, colours", l, L->n_vert, L->n_hex, 2 * L->L_nnz + L->n_vert);
    for (int c = 1; c <= 8; ++c) printf(" %d", L->color_vertex_num[c] - L->color_vertex_num[c - 1]);
    printf(", interpolation entries %d (row sums of Us_hat total %.1f)\n", L->n_entries, wsum);
  }
  femb_setup_hierarchy_free(h);
  free(x); free(hex);
  return 0;
}
