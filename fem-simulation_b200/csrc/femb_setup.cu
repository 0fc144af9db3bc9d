// Host-side setup of the fixed patterns (OpenMP): the integer maps `Object.__init__` builds with per-element Python loops
// and dictionaries (object.py:485-697 under /root/reference/exp1/quasi_simulation) and the symbolic half of the Galerkin
// product that the reference repeats every step inside bsr_mm (sparse.py:1018-1091, object.py:1259).  Same numbering as the
// reference, bit for bit (tests/test_hierarchy.py compares with the literal loops and with the numpy restatement).
//
// Everything here runs on the host on HOST arrays: a level's pattern is "the sorted set of (row, col) over all hexahedra",
// which for a row is the sorted union of the <= 64 index-order ids of the vertices of its <= 8 incident hexahedra -- a
// small per-row sort instead of one global sort of 64 n_hex keys (np.unique: 6 of the 8 s of setup at 1 M hexahedra).
#include <algorithm>
#include <cstring>
#include <vector>
#include <omp.h>

#include "femb_common.cuh"

namespace femb {

// unique index-order columns of row v2i[v], unsorted; returns the count (<= 64).  The <= 64 candidates (8 nodes of each of <= 8
// incident hexahedra) hold ~27 distinct ids: a 128-slot open-addressing table on the stack removes the duplicates before
// anything is sorted
static inline int row_unique(int v, const int* hex, const int* v2i, const int* vert_elem, int* cols) {
  int tab[128];
  std::memset(tab, 0xff, sizeof(tab));
  int n = 0;
  for (int j = 0; j < 8; ++j) {
    const int ea = vert_elem[(size_t)v * 8 + j];
    if (ea < 0) continue;
    const int* h = hex + (size_t)(ea >> 3) * 8;
    for (int b = 0; b < 8; ++b) {
      const int c = v2i[h[b]];
      unsigned s = ((unsigned)c * 2654435761u) >> 25;
      while (tab[s] != -1 && tab[s] != c) s = (s + 1) & 127;
      if (tab[s] == -1) { tab[s] = c; cols[n++] = c; }
    }
  }
  return n;
}

// removes duplicates from cand[0..n) in place (first occurrences kept, unsorted), n <= 224: 512-slot stack table
static inline int unique_in_place(int* cand, int n) {
  int tab[512];
  std::memset(tab, 0xff, sizeof(tab));
  int u = 0;
  for (int k = 0; k < n; ++k) {
    const int c = cand[k];
    unsigned s = ((unsigned)c * 2654435761u) >> 23;
    while (tab[s] != -1 && tab[s] != c) s = (s + 1) & 511;
    if (tab[s] == -1) { tab[s] = c; cand[u++] = c; }
  }
  return u;
}

// sorted unique index-order columns of row v2i[v]; returns the count (<= 64)
static inline int row_columns(int v, const int* hex, const int* v2i, const int* vert_elem, int* cols) {
  const int n = row_unique(v, hex, v2i, vert_elem, cols);
  std::sort(cols, cols + n);
  return n;
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_setup_incidence_host(int n_vert, int n_hex, const int* hex, int* vert_elem) {
  FEMB_REQUIRE(n_vert >= 0 && n_hex >= 0 && hex && vert_elem, "bad arguments");
  std::vector<unsigned char> cnt((size_t)n_vert, 0);
  std::memset(vert_elem, 0xff, sizeof(int) * 8 * (size_t)n_vert);
  for (long long t = 0; t < (long long)n_hex * 8; ++t) {                 // ascending e*8+a: the order of a stable sort by vertex
    const int v = hex[t];
    if (v < 0 || v >= n_vert) { set_error("femb_setup_incidence_host: vertex id %d out of range", v); return FEMB_E_INVALID; }
    if (cnt[v] >= 8) { set_error("femb_setup_incidence_host: vertex %d belongs to more than 8 hexahedra", v); return FEMB_E_INVALID; }
    vert_elem[(size_t)v * 8 + cnt[v]++] = (int)t;
  }
  return FEMB_OK;
}

int femb_setup_coo_to_csr_pair_host(int nrow, int nnz, const int* rows, const int* cols, const float* w, const float* w2, int* offsets,
                                    int* columns, float* weights, float* weights2) {
  FEMB_REQUIRE(nrow >= 0 && nnz >= 0 && offsets && (nnz == 0 || (rows && cols && w && columns && weights)), "bad arguments");
  FEMB_REQUIRE((w2 == nullptr) == (weights2 == nullptr), "w2 and weights2 go together");
  // counting sort by row (stable), then every row by column: entries are unique, rows are short (<= 27 here); a second weight
  // array on the same pattern (the normalised and the un-normalised interpolation weights) rides along
  const bool two = w2 != nullptr;
  // stable counting sort by row, in parallel: thread t owns the t-th contiguous chunk of the input, counts its rows, and the
  // starting position of (row r, thread t) is offsets[r] + the counts of row r in the chunks before t -- input order is kept
  int bad_row = -1;
  const int T = std::max(1, std::min(omp_get_max_threads(), nnz / 65536 + 1));
  std::vector<int> cnt((size_t)T * ((size_t)nrow + 1), 0);
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    const long long lo = (long long)nnz * t / T, hi = (long long)nnz * (t + 1) / T;
    int* c = cnt.data() + (size_t)t * ((size_t)nrow + 1);
    for (long long i = lo; i < hi; ++i) {
      const int r = rows[i];
      if (r < 0 || r >= nrow) {
#pragma omp atomic write
        bad_row = r;
        continue;
      }
      ++c[r];
    }
  }
  if (bad_row != -1) { set_error("femb_setup_coo_to_csr_host: row %d out of range", bad_row); return FEMB_E_INVALID; }
  offsets[0] = 0;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nrow; ++r) {
    int n = 0;
    for (int t = 0; t < T; ++t) {                                         // count -> start of (r, t) relative to the row
      int& c = cnt[(size_t)t * ((size_t)nrow + 1) + r];
      const int k = c;
      c = n; n += k;
    }
    offsets[r + 1] = n;
  }
  for (int r = 0; r < nrow; ++r) offsets[r + 1] += offsets[r];
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    const long long lo = (long long)nnz * t / T, hi = (long long)nnz * (t + 1) / T;
    int* c = cnt.data() + (size_t)t * ((size_t)nrow + 1);
    for (long long i = lo; i < hi; ++i) {
      const int r = rows[i];
      const int p = offsets[r] + c[r]++;
      columns[p] = cols[i]; weights[p] = w[i];
      if (two) weights2[p] = w2[i];
    }
  }
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nrow; ++r)
    for (int i = offsets[r] + 1; i < offsets[r + 1]; ++i) {               // insertion sort, stable
      const int c = columns[i];
      const float x = weights[i], x2 = two ? weights2[i] : 0.f;
      int j = i - 1;
      for (; j >= offsets[r] && columns[j] > c; --j) {
        columns[j + 1] = columns[j]; weights[j + 1] = weights[j];
        if (two) weights2[j + 1] = weights2[j];
      }
      columns[j + 1] = c; weights[j + 1] = x;
      if (two) weights2[j + 1] = x2;
    }
  return FEMB_OK;
}

int femb_setup_coo_to_csr_host(int nrow, int nnz, const int* rows, const int* cols, const float* w, int* offsets, int* columns,
                               float* weights) {
  return femb_setup_coo_to_csr_pair_host(nrow, nnz, rows, cols, w, nullptr, offsets, columns, weights, nullptr);
}

int femb_setup_pattern_count_host(int n_vert, const int* hex, const int* v2i, const int* i2v, const int* vert_elem, int* L_ptr,
                                  int* U_ptr) {
  FEMB_REQUIRE(n_vert >= 0 && hex && v2i && i2v && vert_elem && L_ptr && U_ptr, "bad arguments");
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n_vert; ++r) {
    int cols[64];
    const int n = row_unique(i2v[r], hex, v2i, vert_elem, cols);         // counting needs no order
    int nl = 0, self = 0;
    for (int k = 0; k < n; ++k) { nl += cols[k] < r; self += cols[k] == r; }
    if (n == 0 || !self) bad = 1;                                         // a vertex without a hexahedron has no row
    L_ptr[r + 1] = nl;
    U_ptr[r + 1] = n - nl - 1;
  }
  if (bad) { set_error("femb_setup_pattern_count_host: a vertex belongs to no hexahedron"); return FEMB_E_INVALID; }
  L_ptr[0] = 0; U_ptr[0] = 0;
  for (int r = 0; r < n_vert; ++r) { L_ptr[r + 1] += L_ptr[r]; U_ptr[r + 1] += U_ptr[r]; }
  return FEMB_OK;
}

int femb_setup_pattern_fill_host(int n_vert, int n_hex, const int* hex, const int* v2i, const int* i2v, const int* vert_elem,
                                 const int* L_ptr, const int* U_ptr, int* L_col, int* U_col, int* elem2nnz, int* block_offset) {
  FEMB_REQUIRE(n_vert >= 0 && n_hex >= 0 && hex && v2i && i2v && vert_elem && L_ptr && U_ptr && L_col && U_col && elem2nnz && block_offset,
               "bad arguments");
  const int L_nnz = L_ptr[n_vert];
  const int off_l = 0, off_d = L_nnz, off_u = L_nnz + n_vert;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n_vert; ++r) {
    int cols[64];
    const int n = row_columns(i2v[r], hex, v2i, vert_elem, cols);
    const int nl = L_ptr[r + 1] - L_ptr[r];
    for (int k = 0; k < nl; ++k) L_col[L_ptr[r] + k] = cols[k];
    for (int k = nl + 1; k < n; ++k) U_col[U_ptr[r] + k - nl - 1] = cols[k];
    // rank of the row's blocks in the (row, col)-sorted order of all blocks (UtAUs_block_offset, object.py:633-683)
    const int first = L_ptr[r] + r + U_ptr[r];
    for (int k = 0; k < nl; ++k) block_offset[off_l + L_ptr[r] + k] = first + k;
    block_offset[off_d + r] = first + nl;
    for (int k = nl + 1; k < n; ++k) block_offset[off_u + U_ptr[r] + k - nl - 1] = first + k;
  }
  // hex_update_offset (object.py:563-569): slot of (row of node a, column of node b)
#pragma omp parallel for schedule(static)
  for (int e = 0; e < n_hex; ++e) {
    int idx[8];
    for (int a = 0; a < 8; ++a) idx[a] = v2i[hex[(size_t)e * 8 + a]];
    for (int a = 0; a < 8; ++a) {
      const int r = idx[a];
      const int* lc = L_col + L_ptr[r];
      const int* uc = U_col + U_ptr[r];
      const int nl = L_ptr[r + 1] - L_ptr[r], nu = U_ptr[r + 1] - U_ptr[r];
      for (int b = 0; b < 8; ++b) {
        const int c = idx[b];
        int slot;
        if (c == r) slot = off_d + r;
        else if (c < r) slot = off_l + L_ptr[r] + (int)(std::lower_bound(lc, lc + nl, c) - lc);
        else slot = off_u + U_ptr[r] + (int)(std::lower_bound(uc, uc + nu, c) - uc);
        elem2nnz[(size_t)e * 64 + a * 8 + b] = slot;
      }
    }
  }
  return FEMB_OK;
}

// ---- symbolic Galerkin plan (hierarchy.build_galerkin_plan) ----------------------------------------------------------------
// fine pattern (fL_ptr, fL_col, fU_ptr, fU_col), coarse pattern likewise, Us_hat (fine x coarse) and Ut_hat (coarse x fine) as CSR.
int femb_setup_galerkin_count_host(int n_fine, const int* fL_ptr, const int* fL_col, const int* fU_ptr, const int* fU_col,
                                   const int* us_off, const int* us_col, int* t_ptr) {
  FEMB_REQUIRE(n_fine >= 0 && fL_ptr && fL_col && fU_ptr && fU_col && us_off && us_col && t_ptr, "bad arguments");
  int bad = 0;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_fine; ++i) {
    int cand[27 * 8 + 8];
    int n = 0;
    auto add = [&](int j) { for (int p = us_off[j]; p < us_off[j + 1] && n < 27 * 8 + 8; ++p) cand[n++] = us_col[p]; };
    if (fL_ptr[i + 1] - fL_ptr[i] + 1 + fU_ptr[i + 1] - fU_ptr[i] > 27) { bad = 1; t_ptr[i + 1] = 0; continue; }
    for (int p = fL_ptr[i]; p < fL_ptr[i + 1]; ++p) add(fL_col[p]);
    add(i);
    for (int p = fU_ptr[i]; p < fU_ptr[i + 1]; ++p) add(fU_col[p]);
    const int u = unique_in_place(cand, n);                                // counting needs no order
    if (u > 27) bad = 1;
    t_ptr[i + 1] = u;
  }
  if (bad) { set_error("femb_setup_galerkin_count_host: more than 27 blocks in a row (not a hexahedral lattice hierarchy)"); return FEMB_E_INVALID; }
  t_ptr[0] = 0;
  for (int i = 0; i < n_fine; ++i) t_ptr[i + 1] += t_ptr[i];
  return FEMB_OK;
}

int femb_setup_galerkin_fill_host(int n_fine, int n_coarse, const int* fL_ptr, const int* fL_col, const int* fU_ptr, const int* fU_col,
                                  const int* cL_ptr, const int* cL_col, const int* cU_ptr, const int* cU_col, const int* us_off,
                                  const int* us_col, const float* us_w, const int* ut_off, const int* ut_col, const int* t_ptr,
                                  int* t_col, int* c_ptr, int* c_col, int* c_slot, unsigned char* t_slot, unsigned char* c_slot8,
                                  unsigned char* ut_q, int* dyadic_out) {
  FEMB_REQUIRE(n_fine >= 0 && n_coarse >= 0 && fL_ptr && fU_ptr && cL_ptr && cU_ptr && us_off && us_col && us_w && ut_off && ut_col &&
               t_ptr && t_col && c_ptr && c_col && c_slot && t_slot && c_slot8 && ut_q && dyadic_out, "bad arguments");
  const int fL_nnz = fL_ptr[n_fine], cL_nnz = cL_ptr[n_coarse];
  // (row, col)-sorted pattern of A_c = the coarse hexahedral adjacency, and the [L;D;U] slot of each sorted block
  c_ptr[0] = 0;
  for (int I = 0; I < n_coarse; ++I) c_ptr[I + 1] = c_ptr[I] + (cL_ptr[I + 1] - cL_ptr[I]) + 1 + (cU_ptr[I + 1] - cU_ptr[I]);
  int bad = 0, dyadic = 1;
#pragma omp parallel for schedule(static)
  for (int I = 0; I < n_coarse; ++I) {
    int o = c_ptr[I];
    if (c_ptr[I + 1] - o > 27) bad = 1;
    for (int p = cL_ptr[I]; p < cL_ptr[I + 1]; ++p, ++o) { c_col[o] = cL_col[p]; c_slot[o] = p; }
    c_col[o] = I; c_slot[o] = cL_nnz + I; ++o;
    for (int p = cU_ptr[I]; p < cU_ptr[I + 1]; ++p, ++o) { c_col[o] = cU_col[p]; c_slot[o] = cL_nnz + n_coarse + p; }
  }
  // weights 1, 1/2, 1/4, 1/8 ride as exponents in bits 5-6 of the plan bytes
  const int us_nnz = us_off[n_fine];
#pragma omp parallel for schedule(static) reduction(&& : dyadic)
  for (int p = 0; p < us_nnz; ++p) {
    const float w = us_w[p];
    dyadic = dyadic && (w == 1.0f || w == 0.5f || w == 0.25f || w == 0.125f);
  }
  auto expo = [&](float w) -> unsigned { return w == 1.0f ? 0u : (w == 0.5f ? 1u : (w == 0.25f ? 2u : 3u)); };
  // T rows, and for every fine block (in [L;D;U] slot order) the positions of its column's parents in the T row
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_fine; ++i) {
    int cand[27 * 8 + 8];
    int n = 0;
    auto add = [&](int j) { for (int p = us_off[j]; p < us_off[j + 1] && n < 27 * 8 + 8; ++p) cand[n++] = us_col[p]; };
    for (int p = fL_ptr[i]; p < fL_ptr[i + 1]; ++p) add(fL_col[p]);
    add(i);
    for (int p = fU_ptr[i]; p < fU_ptr[i + 1]; ++p) add(fU_col[p]);
    const int u = unique_in_place(cand, n);                                // ~90 candidates, <= 27 distinct: sort those only
    std::sort(cand, cand + u);
    int* tc = t_col + t_ptr[i];
    for (int k = 0; k < u; ++k) tc[k] = cand[k];
    auto plan = [&](size_t slot, int j) {
      unsigned char* out = t_slot + slot * 8;
      const int np = us_off[j + 1] - us_off[j];
      if (np > 8) { bad = 1; return; }
      for (int p = 0; p < 8; ++p) {
        if (p >= np) { out[p] = 255; continue; }
        const int J = us_col[us_off[j] + p];
        unsigned v = (unsigned)(std::lower_bound(tc, tc + u, J) - tc);
        if (dyadic) v |= expo(us_w[us_off[j] + p]) << 5;
        out[p] = (unsigned char)v;
      }
    };
    for (int p = fL_ptr[i]; p < fL_ptr[i + 1]; ++p) plan((size_t)p, fL_col[p]);
    plan((size_t)fL_nnz + i, i);
    for (int p = fU_ptr[i]; p < fU_ptr[i + 1]; ++p) plan((size_t)fL_nnz + n_fine + p, fU_col[p]);
    // T block t = (i, J): the q-th parent I of i -> position of J in the sorted row I of A_c
    const int npi = us_off[i + 1] - us_off[i];
    for (int k = 0; k < u; ++k) {
      unsigned char* out = c_slot8 + (size_t)(t_ptr[i] + k) * 8;
      for (int q = 0; q < 8; ++q) {
        if (q >= npi) { out[q] = 255; continue; }
        const int I = us_col[us_off[i] + q];
        const int* row = c_col + c_ptr[I];
        const int len = c_ptr[I + 1] - c_ptr[I];
        const int pos = (int)(std::lower_bound(row, row + len, tc[k]) - row);
        if (pos >= len || row[pos] != tc[k]) { bad = 2; out[q] = 255; continue; }
        out[q] = (unsigned char)pos;
      }
    }
  }
  // Ut_hat entry e = (I, i) -> q with parent_q(i) == I
#pragma omp parallel for schedule(static)
  for (int I = 0; I < n_coarse; ++I)
    for (int e = ut_off[I]; e < ut_off[I + 1]; ++e) {
      const int i = ut_col[e];
      const int* par = us_col + us_off[i];
      const int np = us_off[i + 1] - us_off[i];
      const int q = (int)(std::lower_bound(par, par + np, I) - par);
      if (q >= np || par[q] != I) { bad = 3; ut_q[e] = 255; continue; }
      ut_q[e] = (unsigned char)q;
    }
  if (bad) {
    set_error(bad == 2 ? "femb_setup_galerkin_fill_host: the product pattern is not the coarse hexahedral adjacency"
                       : (bad == 3 ? "femb_setup_galerkin_fill_host: Ut_hat is not the transpose of Us_hat"
                                   : "femb_setup_galerkin_fill_host: more than 8 parents or more than 27 blocks in a row"));
    return FEMB_E_INVALID;
  }
  *dyadic_out = dyadic;
  return FEMB_OK;
}

}  // extern "C"
