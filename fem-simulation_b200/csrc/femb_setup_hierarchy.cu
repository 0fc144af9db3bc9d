// Host-side construction of the multigrid hierarchy itself: colour order of a level, one lattice coarsening step (coarse
// hexahedra, coarse vertices, fine_coarse) and the interpolation triplets between two levels -- what Object.__init__ builds
// with Python dictionaries (/root/reference/exp1/quasi_simulation/object.py:147-158, 171-241, 265-443).  The reference numbers
// coarse hexahedra, coarse vertices and interpolation entries in the order its loops first meet them; a dictionary keyed by
// lattice coordinates is a lookup table over the lattice here (dense when the bounding box is reasonably filled, hashed
// otherwise), filled by one ordered scan, so the numbering is the reference's by construction.  Every floating-point
// expression is evaluated in float32 in the reference's order (cpu_function.py:12-20); hierarchy.build_hierarchy_numpy is
// the numpy restatement the tests compare this file with, bit for bit.
//
// Also here, the front end the reference leaves to pyvista / VTK (object.py:28-45): the OBJ parser, the inside test of the
// voxeliser, and voxel set -> points / hexahedra in the reference's numbering (voxelize.py and mesh.py hold the numpy / Python
// restatements), and femb_setup_hierarchy_host, which strings the setup calls of this file and of femb_setup.cu together.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#include <omp.h>

#include "femb_common.cuh"

namespace femb {

// cpu_function.py:12-20: ((p - origin) / spacing + shift).int(), float32 throughout, truncation toward zero
static inline int lattice_index(float p, float origin, float spacing, float shift) {
  float t = p - origin;
  t = t / spacing;
  t = t + shift;
  return (int)t;
}

// first-seen numbering over integer lattice keys (i, j, k) in [0, nx) x [0, ny) x [0, nz)
struct LatticeIds {
  long long nx, ny, nz;
  bool dense;
  std::vector<int> table;
  std::unordered_map<long long, int> map;
  int next = 0;
  LatticeIds(long long nx_, long long ny_, long long nz_, long long expected) : nx(nx_), ny(ny_), nz(nz_) {
    const long long vol = nx * ny * nz;
    dense = vol <= std::max<long long>(1ll << 24, 8 * expected);   // at most 64 MB, or 32 bytes per expected key
    if (dense) table.assign((size_t)vol, -1);
    else map.reserve((size_t)expected * 2);
  }
  inline long long key(long long i, long long j, long long k) const { return (i * ny + j) * nz + k; }
  inline int find_or_add(long long i, long long j, long long k) {
    const long long q = key(i, j, k);
    if (dense) {
      int& slot = table[(size_t)q];
      if (slot < 0) slot = next++;
      return slot;
    }
    auto it = map.find(q);
    if (it != map.end()) return it->second;
    map.emplace(q, next);
    return next++;
  }
};

// object.py:283-412: slot s = 9 i + 3 j + k of a coarse cell (i, j, k in 0..2) -> its coarse corners (ascending local
// number 4 bi + 2 bj + bk) and weight: an even coordinate names one side, the middle one both
struct Stencil {
  int cnt[27];
  int corner[27][8];
  float w[27];
  Stencil() {
    for (int s = 0; s < 27; ++s) {
      const int ax[3] = {s / 9, (s / 3) % 3, s % 3};
      int n = 0;
      for (int c = 0; c < 8; ++c) {
        bool ok = true;
        for (int d = 0; d < 3; ++d) {
          const int bit = (c >> (2 - d)) & 1;
          if (ax[d] != 1 && bit != ax[d] / 2) ok = false;
        }
        if (ok) corner[s][n++] = c;
      }
      cnt[s] = n;
      w[s] = 1.0f / (float)n;
      for (int q = n; q < 8; ++q) corner[s][q] = 0;
    }
  }
};
static const Stencil g_stencil;

// object.py:151-158: index order = stable sort of the vertices by colour
static void color_order(int n, const unsigned char* color, int* v2i, int* i2v, int* cvn) {
  long long count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int v = 0; v < n; ++v) ++count[color[v]];
  long long start[8];
  cvn[0] = 0;
  for (int c = 0; c < 8; ++c) {
    start[c] = cvn[c];
    cvn[c + 1] = (int)(cvn[c] + count[c]);
  }
  for (int v = 0; v < n; ++v) {
    const int i = (int)start[color[v]]++;
    v2i[v] = i;
    i2v[i] = v;
  }
}

template <class T>
static T* grab(size_t n) { return static_cast<T*>(std::malloc(sizeof(T) * std::max<size_t>(n, 1))); }

static inline int mod2(int v) { return ((v % 2) + 2) % 2; }   // numpy's % (result has the divisor's sign)

}  // namespace femb

using namespace femb;

extern "C" {

int femb_setup_stencil_host(int* cnt, int* corners, float* weight) {
  FEMB_REQUIRE(cnt && corners && weight, "null pointer");
  for (int s = 0; s < 27; ++s) {
    cnt[s] = g_stencil.cnt[s];
    weight[s] = g_stencil.w[s];
    for (int q = 0; q < 8; ++q) corners[s * 8 + q] = g_stencil.corner[s][q];
  }
  return FEMB_OK;
}

int femb_setup_color_order_host(int n_vert, const float* x, const float* origin, float spacing, int* v2i, int* i2v,
                                int* color_vertex_num) {
  FEMB_REQUIRE(n_vert >= 0 && x && origin && v2i && i2v && color_vertex_num, "bad arguments");
  std::vector<unsigned char> color((size_t)n_vert);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < n_vert; ++v) {
    const int i = mod2(lattice_index(x[3 * (size_t)v + 0], origin[0], spacing, 0.5f));
    const int j = mod2(lattice_index(x[3 * (size_t)v + 1], origin[1], spacing, 0.5f));
    const int k = mod2(lattice_index(x[3 * (size_t)v + 2], origin[2], spacing, 0.5f));
    color[v] = (unsigned char)(i + 2 * j + 4 * k);
  }
  color_order(n_vert, color.data(), v2i, i2v, color_vertex_num);
  return FEMB_OK;
}

int femb_setup_coarsen_count_host(int n_vert, int n_hex, const float* x, const int* hex, const float* origin, float coarse_size,
                                  int* hex_parent, int* coarse_key, int* n_coarse_hex, int* n_coarse_vert) {
  FEMB_REQUIRE(n_vert > 0 && n_hex > 0 && x && hex && origin && hex_parent && coarse_key && n_coarse_hex && n_coarse_vert,
               "bad arguments");
  // coarse cell of every hexahedron: the lattice index of its node 0 at the coarse spacing (object.py:183-188)
  int kmax[3] = {0, 0, 0};
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : kmax[:3]) reduction(| : bad)
  for (int e = 0; e < n_hex; ++e) {
    const int v = hex[8 * (size_t)e];
    if (v < 0 || v >= n_vert) { bad |= 1; continue; }
    for (int d = 0; d < 3; ++d) {
      const int q = lattice_index(x[3 * (size_t)v + d], origin[d], coarse_size, 0.1f);
      coarse_key[3 * (size_t)e + d] = q;
      if (q < 0 || q >= (1 << 20)) bad |= 2;
      kmax[d] = std::max(kmax[d], q);
    }
  }
  if (bad & 1) { set_error("femb_setup_coarsen_count_host: vertex id out of range"); return FEMB_E_INVALID; }
  if (bad & 2) {
    set_error("femb_setup_coarsen_count_host: coarse cell index outside [0, 2^20): origin above the mesh or spacing too small");
    return FEMB_E_INVALID;
  }
  LatticeIds cells(kmax[0] + 1, kmax[1] + 1, kmax[2] + 1, n_hex);
  for (int e = 0; e < n_hex; ++e)                                          // ordered scan: ids in first-seen order
    hex_parent[e] = cells.find_or_add(coarse_key[3 * (size_t)e], coarse_key[3 * (size_t)e + 1], coarse_key[3 * (size_t)e + 2]);
  const int n_ch = cells.next;
  // keys of the coarse cells in id order, compacted to the front of coarse_key (cell c first appears at a hexahedron >= c)
  {
    std::vector<unsigned char> seen((size_t)n_ch, 0);
    for (int e = 0; e < n_hex; ++e) {
      const int c = hex_parent[e];
      if (seen[c]) continue;
      seen[c] = 1;
      for (int d = 0; d < 3; ++d) coarse_key[3 * (size_t)c + d] = coarse_key[3 * (size_t)e + d];
    }
  }
  LatticeIds verts(kmax[0] + 2, kmax[1] + 2, kmax[2] + 2, (long long)n_ch * 2);
  for (int c = 0; c < n_ch; ++c)
    for (int a = 0; a < 8; ++a)
      verts.find_or_add(coarse_key[3 * (size_t)c] + ((a >> 2) & 1), coarse_key[3 * (size_t)c + 1] + ((a >> 1) & 1),
                        coarse_key[3 * (size_t)c + 2] + (a & 1));
  *n_coarse_hex = n_ch;
  *n_coarse_vert = verts.next;
  return FEMB_OK;
}

int femb_setup_coarsen_fill_host(int n_vert, int n_hex, const float* x, const int* hex, const float* origin, float fine_size,
                                 float coarse_size, const int* hex_parent, const int* coarse_key, int n_coarse_hex,
                                 int n_coarse_vert, int* coarse_hex, float* coarse_x, int* fine_coarse, int* v2i, int* i2v,
                                 int* color_vertex_num) {
  FEMB_REQUIRE(n_vert > 0 && n_hex > 0 && n_coarse_hex > 0 && n_coarse_vert > 0, "empty level");
  FEMB_REQUIRE(x && hex && origin && hex_parent && coarse_key && coarse_hex && coarse_x && fine_coarse && v2i && i2v &&
                   color_vertex_num, "null pointer");
  int kmax[3] = {0, 0, 0};
  for (int c = 0; c < n_coarse_hex; ++c)
    for (int d = 0; d < 3; ++d) {
      const int q = coarse_key[3 * (size_t)c + d];
      FEMB_REQUIRE(q >= 0 && q < (1 << 20), "coarse_key out of range");
      kmax[d] = std::max(kmax[d], q);
    }
  // coarse vertices in the order the corners of the coarse hexahedra first name them (object.py:196-214)
  LatticeIds verts(kmax[0] + 2, kmax[1] + 2, kmax[2] + 2, (long long)n_coarse_hex * 2);
  std::vector<unsigned char> color((size_t)n_coarse_vert);
  for (int c = 0; c < n_coarse_hex; ++c)
    for (int a = 0; a < 8; ++a) {
      const int k3[3] = {coarse_key[3 * (size_t)c] + ((a >> 2) & 1), coarse_key[3 * (size_t)c + 1] + ((a >> 1) & 1),
                         coarse_key[3 * (size_t)c + 2] + (a & 1)};
      const int before = verts.next;
      const int id = verts.find_or_add(k3[0], k3[1], k3[2]);
      if (id >= n_coarse_vert) { set_error("femb_setup_coarsen_fill_host: more coarse vertices than counted"); return FEMB_E_INVALID; }
      coarse_hex[8 * (size_t)c + a] = id;
      if (id == before) {                                                  // new vertex: position and colour (object.py:216-230)
        for (int d = 0; d < 3; ++d) {
          float p = (float)k3[d] * coarse_size;
          p = p + origin[d];
          coarse_x[3 * (size_t)id + d] = p;
        }
        color[id] = (unsigned char)(4 * (k3[2] % 2) + 2 * (k3[1] % 2) + (k3[0] % 2));
      }
    }
  if (verts.next != n_coarse_vert) { set_error("femb_setup_coarsen_fill_host: coarse vertex count differs from the counted one"); return FEMB_E_INVALID; }
  color_order(n_coarse_vert, color.data(), v2i, i2v, color_vertex_num);
  // fine_coarse[c][9 di + 3 dj + dk] = fine vertex at offset (di, dj, dk) in coarse cell c (object.py:189-194, 231-241)
  std::memset(fine_coarse, 0xff, sizeof(int) * 27 * (size_t)n_coarse_hex);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int e = 0; e < n_hex; ++e) {
    const int c = hex_parent[e];
    if (c < 0 || c >= n_coarse_hex) { bad |= 1; continue; }
    for (int b = 0; b < 8; ++b) {
      const int v = hex[8 * (size_t)e + b];
      if (v < 0 || v >= n_vert) { bad |= 1; continue; }
      int slot = 0;
      for (int d = 0; d < 3; ++d) {
        const int q = lattice_index(x[3 * (size_t)v + d], origin[d], fine_size, 0.1f) - 2 * coarse_key[3 * (size_t)c + d];
        if (q < 0 || q > 2) bad |= 2;
        slot = 3 * slot + q;
      }
      if (bad) continue;
      fine_coarse[27 * (size_t)c + slot] = v;                              // every writer of a slot writes the same vertex
    }
  }
  if (bad & 1) { set_error("femb_setup_coarsen_fill_host: index out of range"); return FEMB_E_INVALID; }
  if (bad & 2) { set_error("femb_setup_coarsen_fill_host: a fine vertex lies outside its hexahedron's coarse cell"); return FEMB_E_INVALID; }
  return FEMB_OK;
}

int femb_setup_interp_count_host(int n_fine, int n_coarse_hex, const int* fine_coarse, int* first_pos, int* n_entries) {
  FEMB_REQUIRE(n_fine >= 0 && n_coarse_hex >= 0 && fine_coarse && first_pos && n_entries, "bad arguments");
  FEMB_REQUIRE((long long)n_coarse_hex * 27 < (1ll << 31), "27 n_coarse_hex must fit in int32");
  // the (coarse hexahedron, slot) that first names a fine vertex owns its interpolation row (object.py:265-281)
  for (int v = 0; v < n_fine; ++v) first_pos[v] = -1;
  const int total = n_coarse_hex * 27;
  long long n = 0;
  for (int p = 0; p < total; ++p) {
    const int v = fine_coarse[p];
    if (v == -1) continue;
    if (v < 0 || v >= n_fine) { set_error("femb_setup_interp_count_host: fine vertex %d out of range", v); return FEMB_E_INVALID; }
    if (first_pos[v] < 0) {
      first_pos[v] = p;
      n += g_stencil.cnt[p % 27];
    }
  }
  FEMB_REQUIRE(n < (1ll << 31), "entry count must fit in int32");
  *n_entries = (int)n;
  return FEMB_OK;
}

int femb_setup_interp_fill_host(int n_fine, int n_coarse, int n_coarse_hex, const int* fine_coarse, const int* coarse_hex,
                                const int* first_pos, int n_entries, int* coarse_id, int* fine_id, unsigned char* entry_k,
                                float* weight, float* norm, float* weight_norm) {
  FEMB_REQUIRE(n_fine >= 0 && n_coarse >= 0 && n_coarse_hex >= 0 && n_entries >= 0, "negative size");
  FEMB_REQUIRE(fine_coarse && coarse_hex && first_pos && norm && (n_entries == 0 || (coarse_id && fine_id && entry_k && weight && weight_norm)),
               "null pointer");
  for (int c = 0; c < n_coarse; ++c) norm[c] = 0.0f;
  const int total = n_coarse_hex * 27;
  int n = 0;
  for (int p = 0; p < total; ++p) {                                        // entries in ascending (hexahedron, slot) order
    const int v = fine_coarse[p];
    if (v < 0 || v >= n_fine || first_pos[v] != p) continue;
    const int s = p % 27;
    const int* h = coarse_hex + 8 * (size_t)(p / 27);
    if (n + g_stencil.cnt[s] > n_entries) { set_error("femb_setup_interp_fill_host: more entries than counted"); return FEMB_E_INVALID; }
    for (int k = 0; k < g_stencil.cnt[s]; ++k, ++n) {
      const int c = h[g_stencil.corner[s][k]];
      if (c < 0 || c >= n_coarse) { set_error("femb_setup_interp_fill_host: coarse vertex %d out of range", c); return FEMB_E_INVALID; }
      coarse_id[n] = c;
      fine_id[n] = v;
      entry_k[n] = (unsigned char)k;
      weight[n] = g_stencil.w[s];
      norm[c] += g_stencil.w[s];                                           // dyadic weights: exact in any order (object.py:414-420)
    }
  }
  if (n != n_entries) { set_error("femb_setup_interp_fill_host: entry count differs from the counted one"); return FEMB_E_INVALID; }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_entries; ++i) weight_norm[i] = weight[i] / norm[coarse_id[i]];
  return FEMB_OK;
}


// ---- voxel set -> hexahedral mesh in the reference's numbering (object.py:28-45) -----------------------------------

namespace femb {
// used lattice points of the kept voxels: remap[lattice id] = rank among the used ones (x fastest), -1 if unused
static long long mark_and_rank(int nx, int ny, int nz, const unsigned char* mask, std::vector<int>& remap, long long* n_hex) {
  const long long px = nx + 1, py = ny + 1, pz = nz + 1;
  remap.assign((size_t)(px * py * pz), -1);
  long long nh = 0;
  for (int ix = 0; ix < nx; ++ix)                                            // mask is [nx][ny][nz], z fastest
    for (int iy = 0; iy < ny; ++iy)
      for (int iz = 0; iz < nz; ++iz) {
        if (!mask[((size_t)ix * ny + iy) * nz + iz]) continue;
        ++nh;
        for (int a = 0; a < 8; ++a)
          remap[(size_t)((ix + ((a >> 2) & 1)) + px * ((iy + ((a >> 1) & 1)) + py * (iz + (a & 1))))] = 0;
      }
  long long n = 0;
  for (size_t q = 0; q < remap.size(); ++q)
    if (remap[q] == 0) remap[q] = (int)n++;
  *n_hex = nh;
  return n;
}
}  // namespace femb

int femb_setup_mask_mesh_count_host(int nx, int ny, int nz, const unsigned char* mask, int* n_vert, int* n_hex) {
  FEMB_REQUIRE(nx > 0 && ny > 0 && nz > 0 && mask && n_vert && n_hex, "bad arguments");
  FEMB_REQUIRE((long long)(nx + 1) * (ny + 1) * (nz + 1) < (1ll << 31), "lattice must have fewer than 2^31 points");
  std::vector<int> remap;
  long long nh = 0;
  *n_vert = (int)mark_and_rank(nx, ny, nz, mask, remap, &nh);
  *n_hex = (int)nh;
  return FEMB_OK;
}

int femb_setup_mask_mesh_fill_host(int nx, int ny, int nz, const unsigned char* mask, float dx, const float* origin,
                                   const int* index_offset, int n_vert, int n_hex, float* points, int* hex) {
  FEMB_REQUIRE(nx > 0 && ny > 0 && nz > 0 && mask && origin && index_offset && n_vert >= 0 && n_hex >= 0 && points && hex, "bad arguments");
  FEMB_REQUIRE((long long)(nx + 1) * (ny + 1) * (nz + 1) < (1ll << 31), "lattice must have fewer than 2^31 points");
  std::vector<int> remap;
  long long nh = 0;
  const long long nv = mark_and_rank(nx, ny, nz, mask, remap, &nh);
  if (nv != n_vert || nh != n_hex) { set_error("femb_setup_mask_mesh_fill_host: the mask has %lld vertices / %lld voxels, not %d / %d", nv, nh, n_vert, n_hex); return FEMB_E_INVALID; }
  const long long px = nx + 1, py = ny + 1, pz = nz + 1;
  // points: lattice coordinates (shifted by index_offset) * dx + origin, float32, one multiplication and one addition
#pragma omp parallel for schedule(static)
  for (long long gz = 0; gz < pz; ++gz)
    for (long long gy = 0; gy < py; ++gy)
      for (long long gx = 0; gx < px; ++gx) {
        const int v = remap[(size_t)(gx + px * (gy + py * gz))];
        if (v < 0) continue;
        const long long g[3] = {gx + index_offset[0], gy + index_offset[1], gz + index_offset[2]};
        for (int d = 0; d < 3; ++d) {
          float p = (float)g[d] * dx;
          p = p + origin[d];
          points[3 * (size_t)v + d] = p;
        }
      }
  // hexahedra x fastest, local node a = 4 bx + 2 by + bz
  long long e = 0;
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        if (!mask[((size_t)ix * ny + iy) * nz + iz]) continue;
        for (int a = 0; a < 8; ++a)
          hex[8 * (size_t)e + a] = remap[(size_t)((ix + ((a >> 2) & 1)) + px * ((iy + ((a >> 1) & 1)) + py * (iz + (a & 1))))];
        ++e;
      }
  return FEMB_OK;
}

// ---- Wavefront OBJ surface meshes (object.py:28 calls pv.read): "v x y z" and "f a b c ..." lines, polygons as fans ------------

namespace femb {
// one pass over the file; with verts / faces == nullptr it only counts
static int obj_pass(const char* path, double* verts, long long* faces, long long* n_vert_out, long long* n_tri_out) {
  FILE* f = std::fopen(path, "r");
  if (!f) { set_error("femb_setup_obj: cannot open %s", path); return FEMB_E_INVALID; }
  std::vector<char> line(1 << 16);
  long long nv = 0, nt = 0;
  int rc = FEMB_OK;
  while (std::fgets(line.data(), (int)line.size(), f)) {
    const char* p = line.data();
    if (p[0] == 'v' && p[1] == ' ') {
      double xyz[3] = {0.0, 0.0, 0.0};
      char* end = nullptr;
      const char* q = p + 2;
      for (int d = 0; d < 3; ++d) {
        xyz[d] = std::strtod(q, &end);
        if (end == q) { set_error("femb_setup_obj: malformed vertex line %lld", nv + 1); rc = FEMB_E_INVALID; break; }
        q = end;
      }
      if (rc != FEMB_OK) break;
      if (verts) for (int d = 0; d < 3; ++d) verts[3 * nv + d] = xyz[d];
      ++nv;
    } else if (p[0] == 'f' && p[1] == ' ') {
      long long first = -1, prev = -1;
      int k = 0;
      const char* q = p + 2;
      for (;;) {
        while (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\n') ++q;
        if (!*q) break;
        char* end = nullptr;
        long long i = std::strtoll(q, &end, 10);
        if (end == q) { set_error("femb_setup_obj: malformed face line"); rc = FEMB_E_INVALID; break; }
        i = i > 0 ? i - 1 : nv + i;                                           // 1-based, negative = relative to the vertices so far
        q = end;
        while (*q && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n') ++q;   // skip "/vt/vn"
        if (k == 0) first = i;
        else if (k >= 2) {
          if (faces) { faces[3 * nt] = first; faces[3 * nt + 1] = prev; faces[3 * nt + 2] = i; }
          ++nt;
        }
        prev = i;
        ++k;
      }
      if (rc != FEMB_OK) break;
    }
  }
  std::fclose(f);
  *n_vert_out = nv;
  *n_tri_out = nt;
  return rc;
}
}  // namespace femb

int femb_setup_obj_count_host(const char* path, long long* n_vert, long long* n_tri) {
  FEMB_REQUIRE(path && n_vert && n_tri, "null pointer");
  return obj_pass(path, nullptr, nullptr, n_vert, n_tri);
}

int femb_setup_obj_read_host(const char* path, long long n_vert, long long n_tri, double* verts, long long* faces) {
  FEMB_REQUIRE(path && n_vert >= 0 && n_tri >= 0 && (n_vert == 0 || verts) && (n_tri == 0 || faces), "bad arguments");
  long long nv = 0, nt = 0;
  int rc = obj_pass(path, nullptr, nullptr, &nv, &nt);                        // the file may have changed since the count
  if (rc != FEMB_OK) return rc;
  if (nv != n_vert || nt != n_tri) { set_error("femb_setup_obj_read_host: the file holds %lld vertices / %lld triangles, not %lld / %lld", nv, nt, n_vert, n_tri); return FEMB_E_INVALID; }
  return obj_pass(path, verts, faces, &nv, &nt);
}

// ---- surface mesh -> enclosed lattice points (the inside test of the voxeliser; object.py:28-34 calls pv.voxelize) --------------

int femb_setup_points_inside_host(int n_face, const double* tri, int nx, const double* gx, int ny, const double* gy, int nz,
                                  const double* gz, unsigned char* inside) {
  FEMB_REQUIRE(n_face >= 0 && nx >= 0 && ny >= 0 && nz >= 0 && inside && (n_face == 0 || tri) && gx && gy && gz, "bad arguments");
  for (int i = 1; i < nx; ++i) FEMB_REQUIRE(gx[i] > gx[i - 1], "gx must be strictly increasing");
  for (int i = 1; i < ny; ++i) FEMB_REQUIRE(gy[i] > gy[i - 1], "gy must be strictly increasing");
  for (int i = 1; i < nz; ++i) FEMB_REQUIRE(gz[i] > gz[i - 1], "gz must be strictly increasing");
  std::memset(inside, 0, (size_t)nx * ny * nz);
  // parity of the crossings of the +z ray from every lattice point; a thread owns a range of x indices (a column of the
  // lattice is touched by one thread only), every thread walks all triangles
#pragma omp parallel
  {
    const int T = omp_get_num_threads(), t = omp_get_thread_num();
    const int lo = (int)((long long)nx * t / T), hi = (int)((long long)nx * (t + 1) / T);
    for (int f = 0; f < n_face && lo < hi; ++f) {
      const double* a = tri + 9 * (size_t)f;
      const double *b = a + 3, *c = a + 6;
      const double x0 = std::min(a[0], std::min(b[0], c[0])), x1 = std::max(a[0], std::max(b[0], c[0]));
      const double y0 = std::min(a[1], std::min(b[1], c[1])), y1 = std::max(a[1], std::max(b[1], c[1]));
      const int i0 = std::max(lo, (int)(std::lower_bound(gx, gx + nx, x0) - gx));       // gx >= x0
      const int i1 = std::min(hi, (int)(std::upper_bound(gx, gx + nx, x1) - gx));       // gx <= x1
      const int j0 = (int)(std::lower_bound(gy, gy + ny, y0) - gy);
      const int j1 = (int)(std::upper_bound(gy, gy + ny, y1) - gy);
      if (i0 >= i1 || j0 >= j1) continue;
      const double d = (b[1] - c[1]) * (a[0] - c[0]) + (c[0] - b[0]) * (a[1] - c[1]);
      if (std::fabs(d) < 1e-300) continue;                                              // the ray lies in the triangle's plane
      for (int i = i0; i < i1; ++i)
        for (int j = j0; j < j1; ++j) {
          const double px = gx[i], py = gy[j];
          const double l0 = ((b[1] - c[1]) * (px - c[0]) + (c[0] - b[0]) * (py - c[1])) / d;
          const double l1 = ((c[1] - a[1]) * (px - c[0]) + (a[0] - c[0]) * (py - c[1])) / d;
          const double l2 = 1.0 - l0 - l1;
          // half-open edge rule so that a ray through a shared edge counts once
          const bool hit = d > 0 ? (l0 >= 0 && l1 >= 0 && l2 > 0) : (l0 > 0 && l1 > 0 && l2 >= 0);
          if (!hit) continue;
          const double zc = l0 * a[2] + l1 * b[2] + l2 * c[2];
          unsigned char* col = inside + ((size_t)i * ny + j) * nz;
          for (int k = (int)(std::upper_bound(gz, gz + nz, zc) - gz); k < nz; ++k) col[k] ^= 1;   // gz > zc
        }
    }
  }
  return FEMB_OK;
}

// ---- the whole hierarchy in one call (Object.__init__, object.py:116-697) ----------------------------------------------------

int femb_setup_hierarchy_free(FembHierarchy* h) {
  if (!h) return FEMB_OK;
  for (int l = 0; l < h->n_levels && h->level; ++l) {
    FembHierarchyLevel& L = h->level[l];
    void* owned[] = {L.hex, L.x, L.v2i, L.i2v, L.L_ptr, L.U_ptr, L.L_col, L.U_col, L.elem2nnz, L.block_offset, L.hex_parent,
                     L.fine_coarse, L.coarse_id, L.fine_id, L.entry_k, L.weight, L.weight_norm, L.norm, L.ut_off, L.ut_col,
                     L.ut_w, L.ut_hat_w, L.us_off, L.us_col, L.us_w, L.us_hat_w};
    for (void* q : owned) std::free(q);
  }
  std::free(h->level);
  std::free(h);
  return FEMB_OK;
}

int femb_setup_hierarchy_host(int n_vert, int n_hex, const float* x, const int* hex, float dx, int n_levels, int corrected_spacing,
                              const float* origin, FembHierarchy** out) {
  FEMB_REQUIRE(out, "null output");
  *out = nullptr;
  FEMB_REQUIRE(n_vert > 0 && n_hex > 0 && x && hex && dx > 0.0f && n_levels >= 1 && n_levels <= 32, "bad arguments");
  float o[3];
  for (int d = 0; d < 3; ++d) {
    if (origin) { o[d] = origin[d]; continue; }
    float m = 1000.0f;                                                     // object.py:131-137: box_min starts at 1000
    for (int v = 0; v < n_vert; ++v) m = std::min(m, x[3 * (size_t)v + d]);
    o[d] = m;
  }
  FembHierarchy* H = static_cast<FembHierarchy*>(std::calloc(1, sizeof(FembHierarchy)));
  if (!H) { set_error("femb_setup_hierarchy_host: out of memory"); return FEMB_E_INVALID; }
  H->level = static_cast<FembHierarchyLevel*>(std::calloc((size_t)n_levels, sizeof(FembHierarchyLevel)));
  H->n_levels = n_levels;
  for (int d = 0; d < 3; ++d) H->origin[d] = o[d];
  int rc = FEMB_OK;
#define STEP(call) do { rc = (call); if (rc != FEMB_OK) { femb_setup_hierarchy_free(H); return rc; } } while (0)
#define NEED(ptr) do { if (!(ptr)) { set_error("femb_setup_hierarchy_host: out of memory"); femb_setup_hierarchy_free(H); return FEMB_E_INVALID; } } while (0)
  if (!H->level) { std::free(H); set_error("femb_setup_hierarchy_host: out of memory"); return FEMB_E_INVALID; }
  {                                                                        // level 0: the caller's mesh, copied
    FembHierarchyLevel& L = H->level[0];
    L.n_vert = n_vert; L.n_hex = n_hex; L.spacing = dx;
    NEED(L.hex = grab<int>((size_t)n_hex * 8)); NEED(L.x = grab<float>((size_t)n_vert * 3));
    std::memcpy(L.hex, hex, sizeof(int) * 8 * (size_t)n_hex);
    std::memcpy(L.x, x, sizeof(float) * 3 * (size_t)n_vert);
    NEED(L.v2i = grab<int>(n_vert)); NEED(L.i2v = grab<int>(n_vert));
    STEP(femb_setup_color_order_host(n_vert, L.x, o, dx, L.v2i, L.i2v, L.color_vertex_num));
  }
  for (int l = 0; l + 1 < n_levels; ++l) {                                 // object.py:171-241
    FembHierarchyLevel& F = H->level[l];
    FembHierarchyLevel& Cc = H->level[l + 1];
    float fine_size = l == 0 ? dx : dx * 2.0f;                             // sic (object.py:175-179)
    if (corrected_spacing) fine_size = dx * (float)(1 << l);
    const float coarse_size = fine_size * 2.0f;
    NEED(F.hex_parent = grab<int>(F.n_hex));
    std::vector<int> key((size_t)F.n_hex * 3);
    STEP(femb_setup_coarsen_count_host(F.n_vert, F.n_hex, F.x, F.hex, o, coarse_size, F.hex_parent, key.data(), &Cc.n_hex, &Cc.n_vert));
    Cc.spacing = coarse_size;
    NEED(Cc.hex = grab<int>((size_t)Cc.n_hex * 8)); NEED(Cc.x = grab<float>((size_t)Cc.n_vert * 3));
    NEED(Cc.v2i = grab<int>(Cc.n_vert)); NEED(Cc.i2v = grab<int>(Cc.n_vert));
    NEED(F.fine_coarse = grab<int>((size_t)Cc.n_hex * 27));
    STEP(femb_setup_coarsen_fill_host(F.n_vert, F.n_hex, F.x, F.hex, o, fine_size, coarse_size, F.hex_parent, key.data(), Cc.n_hex,
                                      Cc.n_vert, Cc.hex, Cc.x, F.fine_coarse, Cc.v2i, Cc.i2v, Cc.color_vertex_num));
  }
  for (int l = 0; l + 1 < n_levels; ++l) {                                 // object.py:265-443
    FembHierarchyLevel& F = H->level[l];
    const FembHierarchyLevel& Cc = H->level[l + 1];
    std::vector<int> first_pos((size_t)F.n_vert);
    STEP(femb_setup_interp_count_host(F.n_vert, Cc.n_hex, F.fine_coarse, first_pos.data(), &F.n_entries));
    const size_t ne = (size_t)F.n_entries;
    NEED(F.coarse_id = grab<int>(ne)); NEED(F.fine_id = grab<int>(ne)); NEED(F.entry_k = grab<unsigned char>(ne));
    NEED(F.weight = grab<float>(ne)); NEED(F.weight_norm = grab<float>(ne)); NEED(F.norm = grab<float>(Cc.n_vert));
    STEP(femb_setup_interp_fill_host(F.n_vert, Cc.n_vert, Cc.n_hex, F.fine_coarse, Cc.hex, first_pos.data(), F.n_entries, F.coarse_id,
                                     F.fine_id, F.entry_k, F.weight, F.norm, F.weight_norm));
    // Ut / Ut_hat in index order (object.py:422-436) and their transposes Us / Us_hat (object.py:438-443)
    std::vector<int> row(ne), col(ne);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)ne; ++i) { row[i] = Cc.v2i[F.coarse_id[i]]; col[i] = F.v2i[F.fine_id[i]]; }
    NEED(F.ut_off = grab<int>((size_t)Cc.n_vert + 1)); NEED(F.ut_col = grab<int>(ne));
    NEED(F.ut_w = grab<float>(ne)); NEED(F.ut_hat_w = grab<float>(ne));
    STEP(femb_setup_coo_to_csr_pair_host(Cc.n_vert, F.n_entries, row.data(), col.data(), F.weight_norm, F.weight, F.ut_off, F.ut_col,
                                         F.ut_w, F.ut_hat_w));
    for (int r = 0; r < Cc.n_vert; ++r)
      for (int p = F.ut_off[r]; p < F.ut_off[r + 1]; ++p) row[p] = r;      // rows of the sorted Ut = columns of Us
    NEED(F.us_off = grab<int>((size_t)F.n_vert + 1)); NEED(F.us_col = grab<int>(ne));
    NEED(F.us_w = grab<float>(ne)); NEED(F.us_hat_w = grab<float>(ne));
    STEP(femb_setup_coo_to_csr_pair_host(F.n_vert, F.n_entries, F.ut_col, row.data(), F.ut_w, F.ut_hat_w, F.us_off, F.us_col, F.us_w,
                                         F.us_hat_w));
  }
  for (int l = 0; l < n_levels; ++l) {                                     // object.py:485-697
    FembHierarchyLevel& L = H->level[l];
    std::vector<int> ve((size_t)L.n_vert * 8);
    STEP(femb_setup_incidence_host(L.n_vert, L.n_hex, L.hex, ve.data()));
    NEED(L.L_ptr = grab<int>((size_t)L.n_vert + 1)); NEED(L.U_ptr = grab<int>((size_t)L.n_vert + 1));
    STEP(femb_setup_pattern_count_host(L.n_vert, L.hex, L.v2i, L.i2v, ve.data(), L.L_ptr, L.U_ptr));
    L.L_nnz = L.L_ptr[L.n_vert];
    NEED(L.L_col = grab<int>(L.L_nnz)); NEED(L.U_col = grab<int>(L.L_nnz));
    NEED(L.elem2nnz = grab<int>((size_t)L.n_hex * 64)); NEED(L.block_offset = grab<int>(2 * (size_t)L.L_nnz + L.n_vert));
    STEP(femb_setup_pattern_fill_host(L.n_vert, L.n_hex, L.hex, L.v2i, L.i2v, ve.data(), L.L_ptr, L.U_ptr, L.L_col, L.U_col, L.elem2nnz,
                                      L.block_offset));
  }
#undef STEP
#undef NEED
  *out = H;
  return FEMB_OK;
}

}  // extern "C"
