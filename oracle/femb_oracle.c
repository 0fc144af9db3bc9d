/*
 * ORACLE / CPU BASELINE (test + measurement infrastructure; never linked into the product).
 *
 * Plain-C (OpenMP) restatement of the reference's hot path, kernel by kernel, for (a) the CPU baseline
 * bench.py reports next to the GPU numbers and (b) a second, independent check of the numpy oracle
 * (tests/test_oracle.py compares the two at 1e-5).  The reference is GPU-only NVIDIA-Warp Python; each
 * function below is the loop form of the cited Warp kernel (paths relative to
 * /root/reference/exp1/quasi_simulation/).  Where the Warp kernels scatter with wp.atomic_* (element forces and
 * element Hessian blocks) this baseline does what a CPU implementation would: one pass over the hexahedra into an
 * element scratch, then a gather per vertex / per row through the fixed element->nnz map -- no atomics, fixed summation
 * order -- and the element Hessian uses the closed form of the dF/dE/dP chain (same blocks to round-off; the literal
 * chain lives on in orc_hessian_literal for the cross-check against the numpy oracle).  The per-step COO->BSR rebuilds (bsr_set_from_triplets, a cub
 * sort + host sync per call) and the symbolic half of bsr_mm are NOT reproduced -- patterns are fixed --
 * which only makes this baseline faster than a literal port would be.
 *
 * Parity status: same as oracle/ref_kernels.py (reference ships no golden vectors; wp.eig3 restated).
 *
 * Build:  make -C oracle   ->  oracle/libfemb_oracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* -DORC_DOUBLE builds the float64 twin (same statements): the round-off yardstick of the long-run parity tests */
#ifdef ORC_DOUBLE
typedef double real;
#define SQRT sqrt
#else
typedef float real;
#define SQRT sqrtf
#endif

typedef struct OrcLevel {
  int n_vert, n_hex, nnz, off_l, off_d, off_u;
  int color_start[9];
  const int* hex;       /* [n_hex*8]                        */
  const int* v2i;       /* [n_vert]                         */
  const int* i2v;       /* [n_vert]                         */
  const int* L_ptr;     /* [n_vert+1] rows of self.L        */
  const int* L_col;
  const int* U_ptr;
  const int* U_col;
  const int* elem2nnz;  /* hex_update_offset                */
  const real* invJ;    /* inverse_pX_peps [n_hex*8*9]      */
  const real* detJ;    /* det_pX_peps     [n_hex*8]        */
  const real* mass;    /* m_gpu [n_vert]                   */
  real* values;        /* MF_value_gpu / UtAUs_values      */
  real *B, *B_fixed, *delta_x, *x_solved, *R;   /* dev_B, dev_B_fixed, dev_delta_x, dev_x_solved, dev_R */
  const int* vert_elem; /* [n_vert*8] e*8+a of the hexahedra incident to a vertex, -1 pad (inverse of hex)          */
  real* elem_k;        /* scratch [n_hex*576]: element matrices, block (a,b) at (a*8+b)*9                            */
  real* elem_f;        /* scratch [n_hex*24]: element forces                                                         */
} OrcLevel;

typedef struct OrcCsr {     /* interpolation operator, blocks w*I3 */
  int nrow, ncol, nnz;
  const int* offsets;
  const int* columns;
  const real* weights;
} OrcCsr;

typedef struct OrcPlan {    /* fixed patterns of T = A*Us_hat and A_c (sorted), see hierarchy.build_galerkin_plan */
  int t_nnz;
  const int* t_ptr; const int* t_col;
  const int* c_ptr; const int* c_col; const int* c_slot;
  real* t_val;             /* scratch [t_nnz*9] */
  real* c_val;             /* scratch [nnz_c*9], (row,col)-sorted like self.A[i].values */
} OrcPlan;

typedef struct OrcCtx {
  int layers;
  OrcLevel* lv;
  const OrcCsr* Ut_hat; const OrcCsr* Us_hat; const OrcCsr* Ut_noOrder; const OrcCsr* Ut_noOrder_hat;
  OrcPlan* plan;
  real* x;                  /* x_gpu_layer[0]  [n_vert*3], vertex order */
  real* grad;               /* grad_gpu */
  const int* pin;            /* pin_gpu */
  const real* pin_pos;      /* all_pin_pos_gpu (per vertex) */
  const real* diag_fixed;   /* the k*I pin blocks of MF_value_fixed as a per-index-row diagonal */
  real mu, la, g, control_mag, spd_value;
  real sfg[8][8][3];        /* shapeFuncGrad */
} OrcCtx;

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* launchers such as torchrun export OMP_NUM_THREADS=1; the baseline must say how many threads it uses and set them itself */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------- small dense helpers */
static void mat_mul(const real* a, const real* b, real* c) {      /* c = a b, row-major 3x3 */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static void mat_T(const real* a, real* t) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * j + i];
}

/* F = (sum_i x_i (x) dN_i(q)) * invJ ; E = 0.5 (F^T F - I)            quai_hexagon.py:77-85 */
static void compute_FE(const OrcCtx* c, const OrcLevel* l, const real* x, int e, int q, real* F, real* E) {
  real F0[9];
  for (int r = 0; r < 3; ++r)
    for (int cc = 0; cc < 3; ++cc) {
      real v = 0.f;
      for (int i = 0; i < 8; ++i) v += x[3 * l->hex[8 * e + i] + r] * c->sfg[i][q][cc];
      F0[3 * r + cc] = v;
    }
  mat_mul(F0, l->invJ + ((size_t)e * 8 + q) * 9, F);
  real Ft[9], C[9];
  mat_T(F, Ft);
  mat_mul(Ft, F, C);
  for (int k = 0; k < 9; ++k) E[k] = 0.5f * (C[k] - ((k % 4 == 0) ? 1.f : 0.f));
}

/* compute_partial_elastic_energy_X (quai_hexagon.py:109-133) + gravity (:176-180) + pin springs (:160-165).
 * One pass over the hexahedra into elem_f, then a gather per vertex (incidence order) instead of the kernel's atomics. */
void orc_force(const OrcCtx* c, int level, const real* x, real* grad, int with_gravity_pins) {
  const OrcLevel* l = &c->lv[level];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < l->n_hex; ++e) {
    real fe[24];
    for (int k = 0; k < 24; ++k) fe[k] = 0.f;
    for (int q = 0; q < 8; ++q) {
      real F[9], E[9], S[9], P[9], iT[9], PiT[9];
      compute_FE(c, l, x, e, q, F, E);
      const real tr = E[0] + E[4] + E[8];
      for (int k = 0; k < 9; ++k) S[k] = 2.0f * c->mu * E[k] + ((k % 4 == 0) ? c->la * tr : 0.f);
      mat_mul(F, S, P);
      mat_T(l->invJ + ((size_t)e * 8 + q) * 9, iT);
      mat_mul(P, iT, PiT);
      const real det = l->detJ[(size_t)e * 8 + q];
      for (int i = 0; i < 8; ++i)
        for (int r = 0; r < 3; ++r)
          fe[3 * i + r] -= (PiT[3 * r] * c->sfg[i][q][0] + PiT[3 * r + 1] * c->sfg[i][q][1] + PiT[3 * r + 2] * c->sfg[i][q][2]) * det;
    }
    memcpy(l->elem_f + (size_t)e * 24, fe, sizeof(fe));
  }
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < l->n_vert; ++idx) {
    const int v = l->i2v[idx];
    real f[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < 8; ++j) {
      const int ea = l->vert_elem[(size_t)v * 8 + j];
      if (ea < 0) continue;
      for (int r = 0; r < 3; ++r) f[r] += l->elem_f[(size_t)ea * 3 + r];
    }
    if (with_gravity_pins) {
      f[1] += l->mass[v] * c->g;
      if (c->pin[v] == 1)
        for (int r = 0; r < 3; ++r) f[r] += c->control_mag * (c->pin_pos[3 * v + r] - x[3 * v + r]);
    }
    for (int r = 0; r < 3; ++r) grad[3 * idx + r] = f[r];
  }
}

/* compute_elastic_hessian (quai_hexagon.py:69-106): element matrices by the closed form of the dF/dE/dP chain
 *   K_ab[j][i] = sum_q det ( d_ji (S g_a . g_b) + mu (g_a . g_b)(F F^T)_ji + mu u_b[j] u_a[i] + la u_a[j] u_b[i] ),
 *   g_a = J^-T dN_a, u_a = F g_a, S = 2 mu E + la tr(E) I
 * into elem_k, then every row gathers its blocks through the fixed element->nnz map (hex_update_offset) -- the
 * fixed-pattern, atomics-free assembly a CPU implementation would use; only the row's owner touches the row's slots. */
void orc_hessian(const OrcCtx* c, int level, const real* x, real* values, const real* diag_add) {
  const OrcLevel* l = &c->lv[level];
#pragma omp parallel for schedule(static)
  for (int e = 0; e < l->n_hex; ++e) {
    real Ke[576];
    for (int k = 0; k < 576; ++k) Ke[k] = 0.f;
    for (int q = 0; q < 8; ++q) {
      real F[9], E[9], S[9], M[9], g[8][3], u[8][3], t[8][3];
      compute_FE(c, l, x, e, q, F, E);
      const real* iJ = l->invJ + ((size_t)e * 8 + q) * 9;
      const real det = l->detJ[(size_t)e * 8 + q];
      const real tr = E[0] + E[4] + E[8];
      for (int k = 0; k < 9; ++k) S[k] = 2.0f * c->mu * E[k] + ((k % 4 == 0) ? c->la * tr : 0.f);
      const real cmu = det * c->mu, cla = det * c->la;
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) M[3 * j + i] = cmu * (F[3 * j] * F[3 * i] + F[3 * j + 1] * F[3 * i + 1] + F[3 * j + 2] * F[3 * i + 2]);
      for (int b = 0; b < 8; ++b) {
        for (int j = 0; j < 3; ++j) g[b][j] = c->sfg[b][q][0] * iJ[j] + c->sfg[b][q][1] * iJ[3 + j] + c->sfg[b][q][2] * iJ[6 + j];
        for (int r = 0; r < 3; ++r) u[b][r] = F[3 * r] * g[b][0] + F[3 * r + 1] * g[b][1] + F[3 * r + 2] * g[b][2];
        for (int r = 0; r < 3; ++r) t[b][r] = det * (S[3 * r] * g[b][0] + S[3 * r + 1] * g[b][1] + S[3 * r + 2] * g[b][2]);
      }
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) {
          real* K = Ke + (a * 8 + b) * 9;
          const real s1 = t[a][0] * g[b][0] + t[a][1] * g[b][1] + t[a][2] * g[b][2];
          const real s2 = g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2];
          for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i)
              K[3 * j + i] += M[3 * j + i] * s2 + cmu * u[b][j] * u[a][i] + cla * u[a][j] * u[b][i] + ((i == j) ? s1 : 0.f);
        }
    }
    memcpy(l->elem_k + (size_t)e * 576, Ke, sizeof(Ke));
  }
#pragma omp parallel for schedule(static)
  for (int r = 0; r < l->n_vert; ++r) {
    const int v = l->i2v[r];
    for (int p = l->L_ptr[r]; p < l->L_ptr[r + 1]; ++p) memset(values + ((size_t)l->off_l + p) * 9, 0, 9 * sizeof(real));
    for (int p = l->U_ptr[r]; p < l->U_ptr[r + 1]; ++p) memset(values + ((size_t)l->off_u + p) * 9, 0, 9 * sizeof(real));
    real* d = values + ((size_t)l->off_d + r) * 9;
    memset(d, 0, 9 * sizeof(real));
    if (diag_add) { d[0] = diag_add[r]; d[4] = diag_add[r]; d[8] = diag_add[r]; }
    for (int j = 0; j < 8; ++j) {
      const int ea = l->vert_elem[(size_t)v * 8 + j];
      if (ea < 0) continue;
      const real* Ka = l->elem_k + (size_t)(ea >> 3) * 576 + (size_t)(ea & 7) * 72;
      const int* slot = l->elem2nnz + (size_t)ea * 8;
      for (int b = 0; b < 8; ++b) {
        real* dst = values + (size_t)slot[b] * 9;
        for (int k = 0; k < 9; ++k) dst[k] += Ka[b * 9 + k];
      }
    }
  }
}

/* the same operator as the literal dF/dE/dP chain of quai_hexagon.py:86-106 with the kernel's atomic scatter (one loop
 * iteration per Warp thread): the cross-check of orc_hessian and of the numpy oracle (tests/test_oracle.py) */
void orc_hessian_literal(const OrcCtx* c, int level, const real* x, real* values, const real* diag_add) {
  const OrcLevel* l = &c->lv[level];
  memset(values, 0, sizeof(real) * 9 * (size_t)l->nnz);
  if (diag_add)
    for (int r = 0; r < l->n_vert; ++r)
      for (int k = 0; k < 3; ++k) values[((size_t)l->off_d + r) * 9 + 4 * k] = diag_add[r];
#pragma omp parallel for schedule(static)
  for (long t = 0; t < (long)l->n_hex * 8; ++t) {
    const int e = (int)(t >> 3), q = (int)(t & 7);
    real F[9], E[9], S[9], Ft[9], iT[9];
    compute_FE(c, l, x, e, q, F, E);
    const real* iJ = l->invJ + ((size_t)e * 8 + q) * 9;
    const real det = l->detJ[(size_t)e * 8 + q];
    const real tr = E[0] + E[4] + E[8];
    for (int k = 0; k < 9; ++k) S[k] = 2.0f * c->mu * E[k] + ((k % 4 == 0) ? c->la * tr : 0.f);
    mat_T(F, Ft);
    mat_T(iJ, iT);
    for (int a = 0; a < 8; ++a) {
      const real* dNa = c->sfg[a][q];
      for (int ii = 0; ii < 8; ++ii) {
        real blk[9];
        real gb[3];
        for (int j = 0; j < 3; ++j) gb[j] = c->sfg[ii][q][0] * iJ[j] + c->sfg[ii][q][1] * iJ[3 + j] + c->sfg[ii][q][2] * iJ[6 + j];
        for (int i = 0; i < 3; ++i) {
          real dF[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, dFt[9], t1[9], t2[9], dE[9], dP[9], t3[9], t4[9], M[9];
          for (int j = 0; j < 3; ++j) dF[3 * i + j] = gb[j];
          mat_T(dF, dFt);
          mat_mul(dFt, F, t1);
          mat_mul(Ft, dF, t2);
          for (int k = 0; k < 9; ++k) dE[k] = (t1[k] + t2[k]) * 0.5f;
          const real trd = dE[0] + dE[4] + dE[8];
          mat_mul(dF, S, t3);
          for (int k = 0; k < 9; ++k) M[k] = 2.0f * c->mu * dE[k] + ((k % 4 == 0) ? c->la * trd : 0.f);
          mat_mul(F, M, t4);
          for (int k = 0; k < 9; ++k) dP[k] = t3[k] + t4[k];
          mat_mul(dP, iT, t1);
          for (int j = 0; j < 3; ++j) blk[3 * j + i] = (t1[3 * j] * dNa[0] + t1[3 * j + 1] * dNa[1] + t1[3 * j + 2] * dNa[2]) * det;
        }
        real* dst = values + (size_t)l->elem2nnz[(size_t)e * 64 + a * 8 + ii] * 9;
        for (int k = 0; k < 9; ++k) {
#pragma omp atomic
          dst[k] += blk[k];
        }
      }
    }
  }
}

/* compute_elastic_energy + compute_gravity_energy (quai_hexagon.py:28-64) */
real orc_energy(const OrcCtx* c, const real* x) {
  const OrcLevel* l = &c->lv[0];
  double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
  for (long t = 0; t < (long)l->n_hex * 8; ++t) {
    const int e = (int)(t >> 3), q = (int)(t & 7);
    real F[9], E[9];
    compute_FE(c, l, x, e, q, F, E);
    real s = 0.f;
    for (int k = 0; k < 9; ++k) s += E[k] * E[k];
    const real tr = E[0] + E[4] + E[8];
    tot += (s * c->mu + 0.5f * c->la * tr * tr) * l->detJ[(size_t)e * 8 + q];
  }
#pragma omp parallel for schedule(static) reduction(+ : tot)
  for (int v = 0; v < l->n_vert; ++v) {
    real en = -l->mass[v] * c->g * x[3 * v + 1];
    if (c->pin[v] == 1) {
      real d2 = 0.f;
      for (int r = 0; r < 3; ++r) { const real d = c->pin_pos[3 * v + r] - x[3 * v + r]; d2 += d * d; }
      en += 0.5f * c->control_mag * d2;
    }
    tot += en;
  }
  return (real)tot;
}

/* ---------------------------------------------------------------- sparse algebra (sparse.py:1143-1171) */
static void row_times(const real* vals, const int* cols, int p0, int p1, const real* x, real* acc) {
  for (int p = p0; p < p1; ++p) {
    const real* m = vals + (size_t)p * 9;
    const real* xv = x + 3 * (size_t)cols[p];
    acc[0] += m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2];
    acc[1] += m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2];
    acc[2] += m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2];
  }
}
/* y = alpha * (parts of A) x + beta * y ; parts: 1 L, 2 D, 4 U (three bsr_mv launches in the reference) */
void orc_spmv(const OrcLevel* l, const real* values, int parts, real alpha, const real* x, real beta, real* y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < l->n_vert; ++r) {
    real acc[3] = {0, 0, 0};
    if (parts & 1) row_times(values + (size_t)l->off_l * 9, l->L_col, l->L_ptr[r], l->L_ptr[r + 1], x, acc);
    if (parts & 4) row_times(values + (size_t)l->off_u * 9, l->U_col, l->U_ptr[r], l->U_ptr[r + 1], x, acc);
    if (parts & 2) {
      const real* m = values + ((size_t)l->off_d + r) * 9;
      const real* xv = x + 3 * (size_t)r;
      acc[0] += m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2];
      acc[1] += m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2];
      acc[2] += m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2];
    }
    for (int k = 0; k < 3; ++k) y[3 * r + k] = alpha * acc[k] + (beta != 0.f ? beta * y[3 * r + k] : 0.f);
  }
}

/* solve3x3 (cublas.py:59-91) */
static void solve3x3(const real* A, const real* b, real* x) {
  real r[3] = {b[0], b[1], b[2]}, p[3], Ap[3];
  real r_norm = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (r_norm < 1e-10f) return;
  p[0] = r[0]; p[1] = r[1]; p[2] = r[2];
  for (int l = 0; l < 3; ++l) {
    for (int i = 0; i < 3; ++i) Ap[i] = A[3 * i] * p[0] + A[3 * i + 1] * p[1] + A[3 * i + 2] * p[2];
    const real dot = p[0] * Ap[0] + p[1] * Ap[1] + p[2] * Ap[2];
    if (dot < 1e-10f) return;
    const real alpha = r_norm / dot;
    for (int i = 0; i < 3; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    const real old = r_norm;
    r_norm = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (r_norm < 1e-10f) return;
    const real beta = r_norm / old;
    for (int i = 0; i < 3; ++i) p[i] = r[i] + beta * p[i];
  }
}

/* PerformGaussSeidel (object.py:886-929): B = b - L dx ; dx = 0 ; colours 7..0 {solve ; B[c-1] -= U_c dx} ;
 * B = b - U dx ; dx = 0 ; colours 0..7 {solve ; B[c+1] -= L_c dx} */
void orc_gauss_seidel(OrcLevel* l, int iterations) {
  const real* V = l->values;
  const size_t nb = sizeof(real) * 3 * (size_t)l->n_vert;
  real* dx = l->delta_x; real* B = l->B;
  memset(dx, 0, nb);
  for (int it = 0; it < iterations; ++it) {
    for (int half = 0; half < 2; ++half) {
      /* full product with the "old" triangle */
#pragma omp parallel for schedule(static)
      for (int r = 0; r < l->n_vert; ++r) {
        real acc[3] = {0, 0, 0};
        if (half == 0) row_times(V + (size_t)l->off_l * 9, l->L_col, l->L_ptr[r], l->L_ptr[r + 1], dx, acc);
        else row_times(V + (size_t)l->off_u * 9, l->U_col, l->U_ptr[r], l->U_ptr[r + 1], dx, acc);
        for (int k = 0; k < 3; ++k) B[3 * r + k] = l->B_fixed[3 * r + k] - acc[k];
      }
      memset(dx, 0, nb);
      for (int ci = 0; ci < 8; ++ci) {
        const int c = half == 0 ? 7 - ci : ci;
        const int rb = l->color_start[c], re = l->color_start[c + 1];
#pragma omp parallel for schedule(static)
        for (int r = rb; r < re; ++r) {             /* Colored_GS_MF_Kernel (cublas.py:129-135) */
          real xs[3] = {0, 0, 0};
          solve3x3(V + ((size_t)l->off_d + r) * 9, B + 3 * r, xs);
          dx[3 * r] = xs[0]; dx[3 * r + 1] = xs[1]; dx[3 * r + 2] = xs[2];
        }
        /* slice product onto the next colour's rows (MF_GS_U[c-1] / MF_GS_L[c]) */
        const int cn = half == 0 ? c - 1 : c + 1;
        if (cn < 0 || cn > 7) continue;
        const int nb0 = l->color_start[cn], ne0 = l->color_start[cn + 1];
#pragma omp parallel for schedule(static)
        for (int r = nb0; r < ne0; ++r) {
          real acc[3] = {0, 0, 0};
          if (half == 0) row_times(V + (size_t)l->off_u * 9, l->U_col, l->U_ptr[r], l->U_ptr[r + 1], dx, acc);
          else row_times(V + (size_t)l->off_l * 9, l->L_col, l->L_ptr[r], l->L_ptr[r + 1], dx, acc);
          for (int k = 0; k < 3; ++k) B[3 * r + k] -= acc[k];
        }
      }
    }
  }
}

/* PerformJacobi (object.py:869-884) */
void orc_jacobi(OrcLevel* l, int iterations) {
  const real* V = l->values;
  real* dx = l->delta_x; real* B = l->B;
  memset(dx, 0, sizeof(real) * 3 * (size_t)l->n_vert);
  for (int it = 0; it < iterations; ++it) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < l->n_vert; ++r) {
      real acc[3] = {0, 0, 0};
      row_times(V + (size_t)l->off_l * 9, l->L_col, l->L_ptr[r], l->L_ptr[r + 1], dx, acc);
      row_times(V + (size_t)l->off_u * 9, l->U_col, l->U_ptr[r], l->U_ptr[r + 1], dx, acc);
      for (int k = 0; k < 3; ++k) B[3 * r + k] = l->B_fixed[3 * r + k] - acc[k];
    }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < l->n_vert; ++r) {
      real xs[3] = {0, 0, 0};
      solve3x3(V + ((size_t)l->off_d + r) * 9, B + 3 * r, xs);
      dx[3 * r] = xs[0]; dx[3 * r + 1] = xs[1]; dx[3 * r + 2] = xs[2];
    }
  }
}

void orc_transfer(const OrcCsr* T, real alpha, const real* x, real beta, real* y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < T->nrow; ++r) {
    real s[3] = {0, 0, 0};
    for (int p = T->offsets[r]; p < T->offsets[r + 1]; ++p) {
      const real w = T->weights[p];
      const real* xv = x + 3 * (size_t)T->columns[p];
      s[0] += w * xv[0]; s[1] += w * xv[1]; s[2] += w * xv[2];
    }
    for (int k = 0; k < 3; ++k) y[3 * r + k] = alpha * s[k] + (beta != 0.f ? beta * y[3 * r + k] : 0.f);
  }
}

/* full A x with the three parts (helper for residual / rhs update) */
static void full_mv_sub(const OrcLevel* l, const real* x, real* y) {   /* y -= A x */
  const real* V = l->values;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < l->n_vert; ++r) {
    real acc[3] = {0, 0, 0};
    row_times(V + (size_t)l->off_l * 9, l->L_col, l->L_ptr[r], l->L_ptr[r + 1], x, acc);
    row_times(V + (size_t)l->off_u * 9, l->U_col, l->U_ptr[r], l->U_ptr[r + 1], x, acc);
    const real* m = V + ((size_t)l->off_d + r) * 9;
    const real* xv = x + 3 * (size_t)r;
    acc[0] += m[0] * xv[0] + m[1] * xv[1] + m[2] * xv[2];
    acc[1] += m[3] * xv[0] + m[4] * xv[1] + m[5] * xv[2];
    acc[2] += m[6] * xv[0] + m[7] * xv[1] + m[8] * xv[2];
    for (int k = 0; k < 3; ++k) y[3 * r + k] -= acc[k];
  }
}

/* ---------------------------------------------------------------- spd_matrix33f (cublas.py:115-125), wp.eig3 restated */
static void jacobi_conj(int x, int y, int z, real* s, real* qV) {
  real ch = 2.f * (s[0] - s[2]), sh = s[1];
  const int b = 5.828427124f * sh * sh < ch * ch;
  const real w = 1.0f / SQRT(ch * ch + sh * sh);
  ch = b ? w * ch : 0.923879532f;
  sh = b ? w * sh : 0.3826834323f;
  const real scale = ch * ch + sh * sh, a = (ch * ch - sh * sh) / scale, bb = (2.f * sh * ch) / scale;
  const real t11 = s[0], t21 = s[1], t22 = s[2], t31 = s[3], t32 = s[4], t33 = s[5];
  const real n11 = a * (a * t11 + bb * t21) + bb * (a * t21 + bb * t22);
  const real n21 = a * (-bb * t11 + a * t21) + bb * (-bb * t21 + a * t22);
  const real n22 = -bb * (-bb * t11 + a * t21) + a * (-bb * t21 + a * t22);
  const real n31 = a * t31 + bb * t32, n32 = -bb * t31 + a * t32, n33 = t33;
  real tmp[3] = {qV[0] * sh, qV[1] * sh, qV[2] * sh};
  sh *= qV[3];
  qV[0] *= ch; qV[1] *= ch; qV[2] *= ch; qV[3] *= ch;
  qV[z] += sh; qV[3] -= tmp[z]; qV[x] += tmp[y]; qV[y] -= tmp[x];
  s[0] = n22; s[1] = n32; s[2] = n33; s[3] = n21; s[4] = n31; s[5] = n11;
}
static void spd_block(real* A, real value) {
  real s[6] = {A[0], A[3], A[4], A[6], A[7], A[8]};
  real qV[4] = {0, 0, 0, 1};
  for (int it = 0; it < 4; ++it) { jacobi_conj(0, 1, 2, s, qV); jacobi_conj(1, 2, 0, s, qV); jacobi_conj(2, 0, 1, s, qV); }
  const real x = qV[0], y = qV[1], z = qV[2], w = qV[3];
  real V[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  real d[3];
  for (int k = 0; k < 3; ++k) {
    const real t0 = A[0] * V[k] + A[1] * V[3 + k] + A[2] * V[6 + k];
    const real t1 = A[3] * V[k] + A[4] * V[3 + k] + A[5] * V[6 + k];
    const real t2 = A[6] * V[k] + A[7] * V[3 + k] + A[8] * V[6 + k];
    d[k] = V[k] * t0 + V[3 + k] * t1 + V[6 + k] * t2;
    if (d[k] < 0.f) d[k] = value;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = V[3 * i] * d[0] * V[3 * j] + V[3 * i + 1] * d[1] * V[3 * j + 1] + V[3 * i + 2] * d[2] * V[3 * j + 2];
}
void orc_spd_blocks(real* v, int n, real value) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) spd_block(v + (size_t)i * 9, value);
}

/* ---------------------------------------------------------------- Galerkin: _bsr_mm_compute_values (sparse.py:873-901) twice */
static int lower_bound(const int* a, int lo, int hi, int key) {
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
void orc_galerkin(const OrcCtx* c, int lf) {
  const OrcLevel* f = &c->lv[lf];
  OrcLevel* co = &c->lv[lf + 1];
  const OrcCsr* P = &c->Us_hat[lf];
  const OrcCsr* R = &c->Ut_hat[lf];
  OrcPlan* pl = &c->plan[lf];
  memset(pl->t_val, 0, sizeof(real) * 9 * (size_t)pl->t_nnz);
  /* T = A * Us_hat, thread per row of A; x blocks visited in column order L, D, U */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < f->n_vert; ++i) {
    const int nl = f->L_ptr[i + 1] - f->L_ptr[i], nu = f->U_ptr[i + 1] - f->U_ptr[i];
    for (int jj = 0; jj < nl + 1 + nu; ++jj) {
      int j; const real* a;
      if (jj < nl) { j = f->L_col[f->L_ptr[i] + jj]; a = f->values + ((size_t)f->off_l + f->L_ptr[i] + jj) * 9; }
      else if (jj == nl) { j = i; a = f->values + ((size_t)f->off_d + i) * 9; }
      else { const int u = f->U_ptr[i] + jj - nl - 1; j = f->U_col[u]; a = f->values + ((size_t)f->off_u + u) * 9; }
      for (int p = P->offsets[j]; p < P->offsets[j + 1]; ++p) {
        const int slot = lower_bound(pl->t_col, pl->t_ptr[i], pl->t_ptr[i + 1], P->columns[p]);
        real* t = pl->t_val + (size_t)slot * 9;
        const real w = P->weights[p];
        for (int k = 0; k < 9; ++k) t[k] += a[k] * w;
      }
    }
  }
  const int nnz_c = pl->c_ptr[co->n_vert];
  memset(pl->c_val, 0, sizeof(real) * 9 * (size_t)nnz_c);
#pragma omp parallel for schedule(static)
  for (int I = 0; I < co->n_vert; ++I) {
    for (int e = R->offsets[I]; e < R->offsets[I + 1]; ++e) {
      const int i = R->columns[e];
      const real w = R->weights[e];
      for (int t = pl->t_ptr[i]; t < pl->t_ptr[i + 1]; ++t) {
        const int slot = lower_bound(pl->c_col, pl->c_ptr[I], pl->c_ptr[I + 1], pl->t_col[t]);
        real* d = pl->c_val + (size_t)slot * 9;
        const real* s = pl->t_val + (size_t)t * 9;
        for (int k = 0; k < 9; ++k) d[k] += w * s[k];
      }
    }
  }
  orc_spd_blocks(pl->c_val, nnz_c, c->spd_value);                                     /* object.py:1260 */
#pragma omp parallel for schedule(static)
  for (int s = 0; s < nnz_c; ++s)                                                      /* block_values_reorder, :1261 */
    memcpy(co->values + (size_t)pl->c_slot[s] * 9, pl->c_val + (size_t)s * 9, 9 * sizeof(real));
}

/* ---------------------------------------------------------------- V-cycle and the non-linear steps */
static void axpy(int n, real a, const real* x, real* y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
void orc_vcycle(OrcCtx* c, int layer) {                                                /* object.py:1216-1226 */
  OrcLevel* l = &c->lv[layer];
  const size_t nb = sizeof(real) * 3 * (size_t)l->n_vert;
  orc_gauss_seidel(l, 3);
  if (layer < c->layers - 1) {
    OrcLevel* n = &c->lv[layer + 1];
    memcpy(l->R, l->B_fixed, nb);                                                      /* downSample, :1021-1027 */
    memcpy(l->x_solved, l->delta_x, nb);
    full_mv_sub(l, l->delta_x, l->R);
    orc_transfer(&c->Ut_hat[layer], 1.f, l->R, 0.f, n->B_fixed);
    orc_vcycle(c, layer + 1);
    axpy(3 * n->n_vert, 1.f, n->delta_x, n->x_solved);                                 /* upSample, :1029-1035 */
    orc_transfer(&c->Us_hat[layer], 1.f, n->x_solved, 1.f, l->x_solved);
    full_mv_sub(l, l->x_solved, l->B_fixed);
  }
  if (layer == c->layers - 1) return;
  orc_gauss_seidel(l, 3);
}

static void update_x(const OrcLevel* l, const real* dx, real* x) {                  /* cublas.py:20-24 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < l->n_vert; ++i)
    for (int k = 0; k < 3; ++k) x[3 * l->i2v[i] + k] += dx[3 * i + k];
}

void orc_newton_mg_step(OrcCtx* c) {                                                   /* object.py:1231-1296 */
  OrcLevel* l0 = &c->lv[0];
  orc_force(c, 0, c->x, c->grad, 1);
  orc_hessian(c, 0, c->x, l0->values, c->diag_fixed);
  for (int i = 1; i < c->layers; ++i) orc_galerkin(c, i - 1);
  memcpy(l0->B_fixed, c->grad, sizeof(real) * 3 * (size_t)l0->n_vert);
  orc_vcycle(c, 0);
  axpy(3 * l0->n_vert, 1.f, l0->delta_x, l0->x_solved);                                /* finish, :1037-1040 */
  for (int i = 1; i < c->layers; ++i) memset(c->lv[i].x_solved, 0, sizeof(real) * 3 * (size_t)c->lv[i].n_vert);
  update_x(l0, l0->x_solved, c->x);
}

/* the "newton" loop body of compare() (object.py:1520-1545): assemble, one symmetric GS iteration on level 0, x += dx */
void orc_compare_newton_step(OrcCtx* c) {
  OrcLevel* l0 = &c->lv[0];
  orc_hessian(c, 0, c->x, l0->values, c->diag_fixed);
  orc_force(c, 0, c->x, l0->B_fixed, 1);
  orc_gauss_seidel(l0, 1);
  update_x(l0, l0->delta_x, c->x);
}

void orc_fas_step(OrcCtx* c, int coarse_pin_penalty);
void orc_fasv3_step(OrcCtx* c) { orc_fas_step(c, 1); }                                 /* object.py:1467-1511 */
void orc_fasv2_step(OrcCtx* c) { orc_fas_step(c, 0); }                                 /* object.py:1420-1464: no coarse pin penalty */
real orc_inf_norm(const real* v, int n) {                                              /* Inf_norm, cublas.py:38-48 */
  real m = 0;
  for (int i = 0; i < n; ++i) { const real a = v[i] < 0 ? -v[i] : v[i]; if (a > m) m = a; }
  return m;
}

void orc_fas_step(OrcCtx* c, int coarse_pin_penalty) {
  OrcLevel* l0 = &c->lv[0];
  OrcLevel* l1 = &c->lv[1];
  orc_hessian(c, 0, c->x, l0->values, c->diag_fixed);
  orc_force(c, 0, c->x, l0->B_fixed, 1);
  orc_gauss_seidel(l0, 1);
  update_x(l0, l0->delta_x, c->x);
  orc_transfer(&c->Ut_noOrder[0], 1.f, c->x, 0.f, l1->x_solved);
  orc_force(c, 0, c->x, l0->R, 1);
  orc_transfer(&c->Ut_hat[0], 1.f, l0->R, 0.f, l1->B_fixed);
  orc_hessian(c, 1, l1->x_solved, l1->values, NULL);
  /* compute_fix_hessian (cublas.py:234-242) as a gather over the coarse rows of Ut_noOrder_hat */
  const OrcCsr* T = &c->Ut_noOrder_hat[0];
#pragma omp parallel for schedule(static)
  for (int J = 0; J < (coarse_pin_penalty ? T->nrow : 0); ++J) {
    real s = 0.f;
    for (int p = T->offsets[J]; p < T->offsets[J + 1]; ++p)
      if (c->pin[T->columns[p]] == 1) s += T->weights[p] * T->weights[p] * c->control_mag;
    real* d = l1->values + ((size_t)l1->off_d + l1->v2i[J]) * 9;
    d[0] += s; d[4] += s; d[8] += s;
  }
  orc_jacobi(l1, 1);
  orc_transfer(&c->Us_hat[0], 1.f, l1->delta_x, 0.f, l0->delta_x);
  update_x(l0, l0->delta_x, c->x);
}
