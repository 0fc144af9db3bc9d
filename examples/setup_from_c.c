/* A plain-C consumer of libfemb200's host-side setup: builds the multigrid hierarchy of an n x n x n voxel cube through
 * femb_setup_hierarchy_host (what Object.__init__ of /root/reference/exp1/quasi_simulation/object.py:116-697 computes) and
 * prints one line per level.  No CUDA device is needed for this call; the arrays it returns are what femb_*'s device entry
 * points take after a cudaMemcpy.
 *
 *   gcc -std=c99 -I include examples/setup_from_c.c -L fem-simulation_b200 -lfemb200 -Wl,-rpath,$PWD/fem-simulation_b200 -o setup_from_c
 *   ./setup_from_c 12 3          # cube edge in voxels, number of levels
 *   ./setup_from_c mesh.obj 0.05 2   # a closed surface mesh, voxelised like object.py:28-45 (pv.read + pv.voxelize), 2 levels
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "femb200.h"

static int report(FembHierarchy* h);

/* numpy.arange(lo, hi, dx) in float64: ceil((hi - lo) / dx) values lo + i * ((lo + dx) - lo) */
static double* lattice(double lo, double hi, double dx, int* n) {
  *n = (int)ceil((hi - lo) / dx);
  double* g = (double*)malloc(sizeof(double) * (size_t)(*n > 0 ? *n : 1));
  const double delta = (lo + dx) - lo;
  for (int i = 0; i < *n; ++i) g[i] = lo + (double)i * delta;
  return g;
}

/* OBJ file -> enclosed lattice points -> kept voxels -> mesh -> hierarchy, every stage through the C ABI */
static int from_obj(const char* path, double dx, int levels) {
  long long nv = 0, nt = 0;
  if (femb_setup_obj_count_host(path, &nv, &nt) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  double* verts = (double*)malloc(sizeof(double) * 3 * (size_t)(nv > 0 ? nv : 1));
  long long* faces = (long long*)malloc(sizeof(long long) * 3 * (size_t)(nt > 0 ? nt : 1));
  if (femb_setup_obj_read_host(path, nv, nt, verts, faces) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  double lo[3] = {verts[0], verts[1], verts[2]}, hi[3] = {verts[0], verts[1], verts[2]};
  for (long long v = 0; v < nv; ++v)
    for (int d = 0; d < 3; ++d) {
      if (verts[3 * v + d] < lo[d]) lo[d] = verts[3 * v + d];
      if (verts[3 * v + d] > hi[d]) hi[d] = verts[3 * v + d];
    }
  double* tri = (double*)malloc(sizeof(double) * 9 * (size_t)(nt > 0 ? nt : 1));
  for (long long t = 0; t < nt; ++t)
    for (int c = 0; c < 3; ++c) memcpy(tri + 9 * t + 3 * c, verts + 3 * faces[3 * t + c], sizeof(double) * 3);
  int n[3];
  double* g[3];
  for (int d = 0; d < 3; ++d) g[d] = lattice(lo[d], hi[d], dx, &n[d]);
  unsigned char* inside = (unsigned char*)malloc((size_t)n[0] * n[1] * n[2]);
  if (femb_setup_points_inside_host((int)nt, tri, n[0], g[0], n[1], g[1], n[2], g[2], inside) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  const int cx = n[0] - 1, cy = n[1] - 1, cz = n[2] - 1;                 /* a cell is kept iff one of its corners is enclosed */
  if (cx < 1 || cy < 1 || cz < 1) { fprintf(stderr, "mesh is smaller than one voxel\n"); return 1; }
  unsigned char* mask = (unsigned char*)calloc((size_t)cx * cy * cz, 1);
  for (int ix = 0; ix < cx; ++ix)
    for (int iy = 0; iy < cy; ++iy)
      for (int iz = 0; iz < cz; ++iz)
        for (int a = 0; a < 8; ++a)
          mask[((size_t)ix * cy + iy) * cz + iz] |= inside[((size_t)(ix + ((a >> 2) & 1)) * n[1] + iy + ((a >> 1) & 1)) * n[2] + iz + (a & 1)];
  int n_vert = 0, n_hex = 0;
  if (femb_setup_mask_mesh_count_host(cx, cy, cz, mask, &n_vert, &n_hex) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  float* x = (float*)malloc(sizeof(float) * 3 * (size_t)n_vert);
  int* hex = (int*)malloc(sizeof(int) * 8 * (size_t)n_hex);
  const float origin[3] = {(float)lo[0], (float)lo[1], (float)lo[2]};
  const int offset[3] = {0, 0, 0};
  if (femb_setup_mask_mesh_fill_host(cx, cy, cz, mask, (float)dx, origin, offset, n_vert, n_hex, x, hex) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  printf("obj: %lld vertices, %lld triangles -> %d x %d x %d cells, %d kept\n", nv, nt, cx, cy, cz, n_hex);
  FembHierarchy* h = NULL;
  if (femb_setup_hierarchy_host(n_vert, n_hex, x, hex, (float)dx, levels, 0, NULL, &h) != FEMB_OK) { fprintf(stderr, "%s\n", femb_last_error()); return 1; }
  const int rc = report(h);
  free(verts); free(faces); free(tri); free(inside); free(mask); free(x); free(hex);
  for (int d = 0; d < 3; ++d) free(g[d]);
  return rc;
}

int main(int argc, char** argv) {
  if (argc > 3 && strstr(argv[1], ".obj")) return from_obj(argv[1], atof(argv[2]), atoi(argv[3]));
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
  const int out = report(h);
  free(x); free(hex);
  return out;
}

static int report(FembHierarchy* h) {
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
  return 0;
}
