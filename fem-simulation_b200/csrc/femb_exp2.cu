// exp2 train(): gradient of the two-level step with respect to the interpolation weights w of Ut_noOrder[0].
//
// Reference: /root/reference/exp2/p_hat/quasi_simulation/object.py:1202-1287 -- a wp.Tape around
//   Kronecker_product(w) -> x_solved1 = Ut x_init -> prepare_kernal(level 1) -> x1 = Ut X -> compute_elastic_hessian(level 1)
//   -> jacobi_iteration (solve3x3 on the diagonal blocks) -> Us_hat -> update_deltaX -> force(x) -> loss (atomic_max of |r|)
//   + loss_norm(w)                                                             (cublas.py:93-102, 259-292, train_hexagon.py there)
// followed by Valueaxpy(w, grad, 1e-3).  Warp generates the adjoint kernels; here they are written by hand:
//   * loss: one-hot cotangent at every component that attains the maximum (what the adjoint of atomic_max hands back);
//   * force: its Jacobian is minus the fine Hessian (symmetric), so the cotangent of the state is one femb_hessian +
//     femb_spmv by the caller; Us_hat^T is Ut_hat (femb_transfer_mv);
//   * solve3x3 and the coarse diagonal blocks (functions of x1 AND of the coarse rest shape, i.e. third derivatives of the
//     SVK energy and the derivative of det / inverse of the rest Jacobian): vector-Jacobian products by FORWARD mode on dual
//     numbers -- the primal code instantiated on (value, tangent) pairs, one thread per input direction.  The cotangents are
//     nonzero on a handful of coarse vertices only (the loss is a max), so the extra arithmetic is irrelevant, and the primal
//     code is shared with the forward path instead of being transposed by hand.
#include "femb_common.cuh"

namespace femb {
namespace x2t {

struct Dual {
  float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d = 0.f) { Dual r; r.v = v; r.d = d; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return mk(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.v * b.d + a.d * b.v); }
__device__ __forceinline__ Dual operator*(float a, Dual b) { return mk(a * b.v, a * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return mk(q, (a.d - q * b.d) / b.v);
}

// solve3x3 (cublas.py:59-91) from x = 0 on dual numbers: the branches follow the values, the tangents ride along
__device__ __forceinline__ void solve3x3_dual(const Dual (&A)[9], const float (&b)[3], Dual (&x)[3]) {
  x[0] = x[1] = x[2] = mk(0.f);
  Dual r[3] = {mk(b[0]), mk(b[1]), mk(b[2])};
  Dual rn = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (rn.v < 1e-10f) return;
  Dual p[3] = {r[0], r[1], r[2]};
#pragma unroll 1
  for (int l = 0; l < 3; ++l) {
    Dual Ap[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Ap[i] = A[3 * i] * p[0] + A[3 * i + 1] * p[1] + A[3 * i + 2] * p[2];
    const Dual dot = p[0] * Ap[0] + p[1] * Ap[1] + p[2] * Ap[2];
    if (dot.v < 1e-10f) return;
    const Dual alpha = rn / dot;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      x[i] = x[i] + alpha * p[i];
      r[i] = r[i] - alpha * Ap[i];
    }
    const Dual old = rn;
    rn = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (rn.v < 1e-10f) return;
    const Dual beta = rn / old;
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = r[i] + beta * p[i];
  }
}

// Dbar[row][k] = lam[row] . d solve3x3(D[row], b[row]) / d D[row][k]: thread (row, k)
__global__ void __launch_bounds__(256) solve3x3_vjp_kernel(int n, const float* __restrict__ D, const float* __restrict__ b,
                                                          const float* __restrict__ lam, float* __restrict__ Dbar) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 9) return;
  const int row = t / 9, k = t - row * 9;
  const float l0 = lam[3 * (size_t)row], l1 = lam[3 * (size_t)row + 1], l2 = lam[3 * (size_t)row + 2];
  float out = 0.f;
  if (l0 != 0.f || l1 != 0.f || l2 != 0.f) {
    Dual A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = mk(D[(size_t)row * 9 + i], i == k ? 1.f : 0.f);
    const float bb[3] = {b[3 * (size_t)row], b[3 * (size_t)row + 1], b[3 * (size_t)row + 2]};
    Dual x[3];
    solve3x3_dual(A, bb, x);
    out = l0 * x[0].d + l1 * x[1].d + l2 * x[2].d;
  }
  Dbar[t] = out;
}

// onehot[i] = sign(v[i]) where |v[i]| attains the maximum (maxv[0]), else 0: the cotangent atomic_max hands back
__global__ void __launch_bounds__(256) absmax_onehot_kernel(int n, const float* __restrict__ v, const float* __restrict__ maxv,
                                                           float* __restrict__ onehot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = v[i];
  onehot[i] = fabsf(x) == maxv[0] ? (x > 0.f ? 1.f : -1.f) : 0.f;
}

__device__ __forceinline__ float sfg_rt(int a, int q, int c) {
  const float s = ((a >> (2 - c)) & 1) ? 1.0f : -1.0f;
  const int d0 = (c == 0) ? 1 : 0, d1 = (c == 2) ? 1 : 2;
  const float f0 = (((a >> (2 - d0)) & 1) == ((q >> (2 - d0)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
  const float f1 = (((a >> (2 - d1)) & 1) == ((q >> (2 - d1)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
  return (s * f0) * f1 * 0.125f;
}

// G[e][dir] = sum_a Dbar[hex[e][a]] : d K_aa / d input_dir, K_aa = the diagonal block of node a of the element matrix of coarse
// hexahedron e (prepare_kernal + compute_mat_F + compute_elastic_hessian of train_hexagon.py / temp_kernel.py as ONE function
// of the 8 node positions X and the 8 rest positions R).  dir < 24: X[dir/3][dir%3]; dir >= 24: R[(dir-24)/3][(dir-24)%3].
constexpr int DIRS = 48;
__global__ void __launch_bounds__(DIRS * 2) coarse_diag_vjp_kernel(int n_hex, const int* __restrict__ hex, const float* __restrict__ x,
                                                                  const float* __restrict__ xrest, float mu, float la,
                                                                  const float* __restrict__ Dbar_vertex, float* __restrict__ G) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_hex * DIRS) return;
  const int e = t / DIRS, dir = t - e * DIRS;
  int vid[8];
  bool any = false;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    vid[a] = __ldg(hex + (size_t)e * 8 + a);
#pragma unroll
    for (int k = 0; k < 9; ++k) any = any || (Dbar_vertex[(size_t)vid[a] * 9 + k] != 0.f);
  }
  if (!any) { G[t] = 0.f; return; }
  Dual X[8][3], R[8][3];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      X[a][c] = mk(x[3 * (size_t)vid[a] + c], dir == 3 * a + c ? 1.f : 0.f);
      R[a][c] = mk(xrest[3 * (size_t)vid[a] + c], dir == 24 + 3 * a + c ? 1.f : 0.f);
    }
  float acc = 0.f;
#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    float dN[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) dN[a][c] = sfg_rt(a, q, c);
    // rest Jacobian, its determinant and inverse (prepare_kernal)
    Dual J[9], F0[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Dual sj = mk(0.f), sf = mk(0.f);
#pragma unroll
        for (int a = 0; a < 8; ++a) { sj = sj + dN[a][c] * R[a][r]; sf = sf + dN[a][c] * X[a][r]; }
        J[3 * r + c] = sj; F0[3 * r + c] = sf;
      }
    const Dual c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    const Dual det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const Dual id = mk(1.f) / det;
    Dual iJ[9];
    iJ[0] = c00 * id; iJ[1] = (J[2] * J[7] - J[1] * J[8]) * id; iJ[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    iJ[3] = c01 * id; iJ[4] = (J[0] * J[8] - J[2] * J[6]) * id; iJ[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    iJ[6] = c02 * id; iJ[7] = (J[1] * J[6] - J[0] * J[7]) * id; iJ[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    Dual F[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) F[3 * r + c] = F0[3 * r] * iJ[c] + F0[3 * r + 1] * iJ[3 + c] + F0[3 * r + 2] * iJ[6 + c];
    // E = 0.5 (F^T F - I), S = 2 mu E + la tr(E) I, FF^T   (symmetric: 00 01 02 11 12 22)
    Dual E[6];
    E[0] = 0.5f * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6] - mk(1.f));
    E[1] = 0.5f * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
    E[2] = 0.5f * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]);
    E[3] = 0.5f * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7] - mk(1.f));
    E[4] = 0.5f * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
    E[5] = 0.5f * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8] - mk(1.f));
    const Dual trE = E[0] + E[3] + E[5];
    const float tm = 2.f * mu;
    Dual S[6];
    S[0] = tm * E[0] + la * trE; S[1] = tm * E[1]; S[2] = tm * E[2];
    S[3] = tm * E[3] + la * trE; S[4] = tm * E[4]; S[5] = tm * E[5] + la * trE;
    Dual FFT[6];
    FFT[0] = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
    FFT[1] = F[0] * F[3] + F[1] * F[4] + F[2] * F[5];
    FFT[2] = F[0] * F[6] + F[1] * F[7] + F[2] * F[8];
    FFT[3] = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
    FFT[4] = F[3] * F[6] + F[4] * F[7] + F[5] * F[8];
    FFT[5] = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
    const Dual cmu = mu * det, cuu = (mu + la) * det;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float* db = Dbar_vertex + (size_t)vid[a] * 9;
      Dual g[3], u[3], tt[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) g[j] = dN[a][0] * iJ[j] + dN[a][1] * iJ[3 + j] + dN[a][2] * iJ[6 + j];
#pragma unroll
      for (int r = 0; r < 3; ++r) u[r] = F[3 * r] * g[0] + F[3 * r + 1] * g[1] + F[3 * r + 2] * g[2];
      tt[0] = det * (S[0] * g[0] + S[1] * g[1] + S[2] * g[2]);
      tt[1] = det * (S[1] * g[0] + S[3] * g[1] + S[4] * g[2]);
      tt[2] = det * (S[2] * g[0] + S[4] * g[1] + S[5] * g[2]);
      const Dual s1 = tt[0] * g[0] + tt[1] * g[1] + tt[2] * g[2];
      const Dual s2 = cmu * (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
      const Dual* Fs[9] = {&FFT[0], &FFT[1], &FFT[2], &FFT[1], &FFT[3], &FFT[4], &FFT[2], &FFT[4], &FFT[5]};
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          Dual kji = s2 * (*Fs[3 * j + i]) + cuu * (u[j] * u[i]);
          if (i == j) kji = kji + s1;
          acc += db[3 * j + i] * kji.d;
        }
    }
  }
  G[t] = acc;
}

// lam_x[J] = sum over the hexahedra around coarse vertex J of G[e][3 a + c], lam_rest[J] likewise from G[e][24 + 3 a + c]
__global__ void __launch_bounds__(256) gather_dirs_kernel(int n_vert, const int* __restrict__ vert_elem, const float* __restrict__ G,
                                                         float* __restrict__ lam_x, float* __restrict__ lam_rest) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  float sx[3] = {0.f, 0.f, 0.f}, sr[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < 8; ++j) {
    const int ea = __ldg(vert_elem + (size_t)v * 8 + j);
    if (ea < 0) continue;
    const float* g = G + (size_t)(ea >> 3) * DIRS + 3 * (ea & 7);
#pragma unroll
    for (int c = 0; c < 3; ++c) { sx[c] += g[c]; sr[c] += g[24 + c]; }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { lam_x[3 * (size_t)v + c] = sx[c]; lam_rest[3 * (size_t)v + c] = sr[c]; }
}

// grad[p] for entry p = (J, i) of Ut_noOrder: lam_x[J] . X[i] + lam_rest[J] . x_init[i] + 2 alpha (1 - rowsum_J) (loss_norm,
// rows with |1 - rowsum| >= 1e-6 only); thread per coarse row
__global__ void __launch_bounds__(256) weight_grad_kernel(int n_row, const int* __restrict__ off, const int* __restrict__ col,
                                                         const float* __restrict__ w, const float* __restrict__ lam_x,
                                                         const float* __restrict__ X, const float* __restrict__ lam_rest,
                                                         const float* __restrict__ x_init, float alpha, float* __restrict__ grad,
                                                         float* __restrict__ reg_rows) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  if (J >= n_row) return;
  const int p0 = __ldg(off + J), p1 = __ldg(off + J + 1);
  float s = 0.f;
  for (int p = p0; p < p1; ++p) s += w[p];
  const float dev = 1.0f - s;
  const bool live = fabsf(dev) >= 1e-6f;
  if (reg_rows != nullptr) reg_rows[J] = live ? dev * dev : 0.f;
  const float lx[3] = {lam_x[3 * (size_t)J], lam_x[3 * (size_t)J + 1], lam_x[3 * (size_t)J + 2]};
  const float lr[3] = {lam_rest[3 * (size_t)J], lam_rest[3 * (size_t)J + 1], lam_rest[3 * (size_t)J + 2]};
  for (int p = p0; p < p1; ++p) {
    const int i = __ldg(col + p);
    float g = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) g += lx[c] * X[3 * (size_t)i + c] + lr[c] * x_init[3 * (size_t)i + c];
    grad[p] = g + (live ? 2.f * alpha * dev : 0.f);
  }
}

// Valueaxpy (cublas.py:156-166): w = clamp(w + a g, 0, 1)
__global__ void __launch_bounds__(256) weight_update_kernel(int n, float* __restrict__ w, const float* __restrict__ g, float a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float y = w[i] + a * g[i];
  w[i] = y < 0.f ? 0.f : (y > 1.f ? 1.f : y);
}

}  // namespace x2t
}  // namespace femb

using namespace femb;

extern "C" {

int femb_absmax_onehot(int n, const float* v, const float* maxv, float* onehot, void* stream) {
  FEMB_REQUIRE(n >= 0 && v && maxv && onehot, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  x2t::absmax_onehot_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, v, maxv, onehot);
  count_launch();
  return check_launch("absmax_onehot_kernel");
}

int femb_solve3x3_vjp(int n, const float* diag_blocks, const float* b, const float* lam, float* dbar, void* stream) {
  FEMB_REQUIRE(n >= 0 && diag_blocks && b && lam && dbar, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  x2t::solve3x3_vjp_kernel<<<cdiv((long long)n * 9, 256), 256, 0, as_stream(stream)>>>(n, diag_blocks, b, lam, dbar);
  count_launch();
  return check_launch("solve3x3_vjp_kernel");
}

int femb_coarse_diag_vjp(const FembLevel* lvl, const float* x, const float* x_rest, float mu, float lambda, const float* dbar_vertex,
                         float* elem_scratch, float* lam_x, float* lam_rest, void* stream) {
  FEMB_REQUIRE(lvl && x && x_rest && dbar_vertex && elem_scratch && lam_x && lam_rest, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->vert_elem, "level not initialised");
  cudaStream_t st = as_stream(stream);
  if (lvl->n_hex > 0) {
    x2t::coarse_diag_vjp_kernel<<<cdiv((long long)lvl->n_hex * x2t::DIRS, x2t::DIRS * 2), x2t::DIRS * 2, 0, st>>>(
        lvl->n_hex, lvl->hex, x, x_rest, mu, lambda, dbar_vertex, elem_scratch);
    count_launch();
  }
  if (lvl->n_vert > 0) {
    x2t::gather_dirs_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, st>>>(lvl->n_vert, lvl->vert_elem, elem_scratch, lam_x, lam_rest);
    count_launch();
  }
  return check_launch("femb_coarse_diag_vjp");
}

int femb_exp2_weight_grad(const FembTransfer* ut, const float* w, const float* lam_x, const float* X, const float* lam_rest,
                          const float* x_init, float alpha, float* grad, float* reg_rows, void* stream) {
  FEMB_REQUIRE(ut && w && lam_x && X && lam_rest && x_init && grad, "null pointer");
  if (ut->nrow == 0) return FEMB_OK;
  x2t::weight_grad_kernel<<<cdiv(ut->nrow, 256), 256, 0, as_stream(stream)>>>(ut->nrow, ut->offsets, ut->columns, w, lam_x, X, lam_rest,
                                                                                x_init, alpha, grad, reg_rows);
  count_launch();
  return check_launch("weight_grad_kernel");
}

int femb_exp2_weight_update(int n, float* w, const float* grad, float lr, void* stream) {
  FEMB_REQUIRE(n >= 0 && w && grad, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  x2t::weight_update_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, w, grad, lr);
  count_launch();
  return check_launch("weight_update_kernel");
}

}  // extern "C"
