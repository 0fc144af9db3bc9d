// Sparse algebra on the fixed 3x3-block [L;D;U] pattern: SpMV, residual, 8-colour symmetric block
// Gauss-Seidel, block Jacobi, CG, grid transfers, Galerkin triple product, spd_matrix33f.
//
// Reference (paths relative to /root/reference/exp1/quasi_simulation/):
//   _bsr_mv_kernel            sparse.py:1143-1171  (thread per row, 36-byte uncoalesced block loads)
//   PerformGaussSeidel        object.py:886-929    (per iteration: 2 full bsr_mv + 16 colour kernels +
//                                                   14 bsr_set_from_triplets (sort + host sync) + 14 slice bsr_mv)
//   PerformJacobi             object.py:869-884
//   PerformConjugateGradient  object.py:931-1019   (2 host read-backs per iteration)
//   bsr_mm                    sparse.py:822-1140   (symbolic + numeric every step)
//
// B200 design.  All of the above are one "row pass": for a contiguous range of index-order rows the
// CTA streams the rows' L blocks, U blocks (both contiguous in the [L;D;U] layout) and D blocks from
// HBM into shared memory with fully coalesced loads, then 4 threads per row walk the staged blocks
// (lanes = consecutive rows, so the x[col] gathers of a lattice mesh coalesce), and an epilogue turns
// the row sum into y / residual / a solve3x3 update.  A colour of the 8-colouring is a contiguous row
// range, so one symmetric Gauss-Seidel iteration is 16 such launches that update dx in place: rows of
// colour c read neighbours of lower colours (old values in the backward sweep, new in the forward
// sweep) and higher colours (the opposite) -- exactly b - L dx_old - U dx_new of object.py:890-908
// without the per-colour matrix rebuilds.  Matrix bytes move once per half-sweep.
#include "femb_common.cuh"

namespace femb {

constexpr int RP_ROWS = 32;                 // rows per CTA
constexpr int RP_SUBS = 4;                  // threads per row
constexpr int RP_THREADS = RP_ROWS * RP_SUBS;
constexpr int RP_CAP = RP_ROWS * 27;        // staged off-diagonal blocks per CTA (26 neighbours + slack)

enum RowMode { RM_SPMV = 0, RM_GS = 1, RM_JACOBI = 2, RM_RESIDUAL = 3 };

struct RowArgs {
  const int* L_ptr; const int* L_col; const int* U_ptr; const int* U_col;
  const float* values;
  int off_l, off_d, off_u;
  int row_begin, row_end;      // rows handled by this launch
  int parts;                   // FEMB_PART_* mask of the matrix parts that enter the row sum
  const float* x;              // vector multiplied by the off-diagonal/diagonal blocks
  const float* b;              // right-hand side (GS/JACOBI/RESIDUAL)
  float* y;                    // output
  float alpha, beta;           // SPMV: y = alpha*sum + beta*y
  int* not_spd;
  const float* active;         // optional device flag: 0 -> kernel is a no-op (CG early exit)
};

template <int kMode>
__global__ void __launch_bounds__(RP_THREADS) row_pass_kernel(RowArgs a) {
  if (a.active != nullptr && *a.active == 0.f) return;
  __shared__ float s_val[RP_CAP * 9];
  __shared__ int s_col[RP_CAP];
  __shared__ float s_diag[RP_ROWS * 9];
  __shared__ int s_ptr[2][RP_ROWS + 1];
  __shared__ float s_red[RP_SUBS][RP_ROWS][3];
  const int tid = threadIdx.x;
  const int r0 = a.row_begin + blockIdx.x * RP_ROWS;
  const int nrow = min(RP_ROWS, a.row_end - r0);
  const bool useL = (a.parts & FEMB_PART_L) != 0, useU = (a.parts & FEMB_PART_U) != 0;
  const bool needD = (kMode == RM_GS || kMode == RM_JACOBI) || (a.parts & FEMB_PART_D) != 0;
  if (tid <= nrow) {
    s_ptr[0][tid] = useL ? __ldg(a.L_ptr + r0 + tid) : 0;
    s_ptr[1][tid] = useU ? __ldg(a.U_ptr + r0 + tid) : 0;
  }
  __syncthreads();
  const int L0 = s_ptr[0][0], U0 = s_ptr[1][0];
  const int nL = s_ptr[0][nrow] - L0, nU = s_ptr[1][nrow] - U0;
  const bool staged = (nL + nU) <= RP_CAP;
  if (staged) {
    const float* srcL = a.values + ((size_t)a.off_l + L0) * 9;
    for (int i = tid; i < nL * 9; i += RP_THREADS) s_val[i] = __ldg(srcL + i);
    const float* srcU = a.values + ((size_t)a.off_u + U0) * 9;
    float* dstU = s_val + nL * 9;
    for (int i = tid; i < nU * 9; i += RP_THREADS) dstU[i] = __ldg(srcU + i);
    for (int i = tid; i < nL; i += RP_THREADS) s_col[i] = __ldg(a.L_col + L0 + i);
    for (int i = tid; i < nU; i += RP_THREADS) s_col[nL + i] = __ldg(a.U_col + U0 + i);
  }
  if (needD) {
    const float* srcD = a.values + ((size_t)a.off_d + r0) * 9;
    for (int i = tid; i < nrow * 9; i += RP_THREADS) s_diag[i] = __ldg(srcD + i);
  }
  __syncthreads();

  const int rl = tid % RP_ROWS, sub = tid / RP_ROWS;
  float acc[3] = {0.f, 0.f, 0.f};
  if (rl < nrow) {
    const int lb = s_ptr[0][rl] - L0, cntL = s_ptr[0][rl + 1] - s_ptr[0][rl];
    const int ub = s_ptr[1][rl] - U0, cntU = s_ptr[1][rl + 1] - s_ptr[1][rl];
    const int n = cntL + cntU;
    for (int j = sub; j < n; j += RP_SUBS) {
      const bool isL = j < cntL;
      const int loc = isL ? (lb + j) : (nL + ub + (j - cntL));
      int col;
      float m[9];
      if (staged) {
        col = s_col[loc];
#pragma unroll
        for (int k = 0; k < 9; ++k) m[k] = s_val[loc * 9 + k];
      } else {
        const size_t g = isL ? ((size_t)a.off_l + L0 + lb + j) : ((size_t)a.off_u + U0 + ub + (j - cntL));
        col = isL ? __ldg(a.L_col + L0 + lb + j) : __ldg(a.U_col + U0 + ub + (j - cntL));
#pragma unroll
        for (int k = 0; k < 9; ++k) m[k] = __ldg(a.values + g * 9 + k);
      }
      const float x0 = a.x[3 * (size_t)col], x1 = a.x[3 * (size_t)col + 1], x2 = a.x[3 * (size_t)col + 2];
      acc[0] += m[0] * x0 + m[1] * x1 + m[2] * x2;
      acc[1] += m[3] * x0 + m[4] * x1 + m[5] * x2;
      acc[2] += m[6] * x0 + m[7] * x1 + m[8] * x2;
    }
  }
  s_red[sub][rl][0] = acc[0]; s_red[sub][rl][1] = acc[1]; s_red[sub][rl][2] = acc[2];
  __syncthreads();
  if (tid >= nrow) return;
  const int r = r0 + tid;
  float s[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) s[c] = (s_red[0][tid][c] + s_red[1][tid][c]) + (s_red[2][tid][c] + s_red[3][tid][c]);
  const float* D = s_diag + tid * 9;
  if (kMode == RM_SPMV || kMode == RM_RESIDUAL) {
    if (a.parts & FEMB_PART_D) {
      const float x0 = a.x[3 * (size_t)r], x1 = a.x[3 * (size_t)r + 1], x2 = a.x[3 * (size_t)r + 2];
      s[0] += D[0] * x0 + D[1] * x1 + D[2] * x2;
      s[1] += D[3] * x0 + D[4] * x1 + D[5] * x2;
      s[2] += D[6] * x0 + D[7] * x1 + D[8] * x2;
    }
  }
  if (kMode == RM_SPMV) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = a.alpha * s[c];
      if (a.beta != 0.f) v += a.beta * a.y[3 * (size_t)r + c];
      a.y[3 * (size_t)r + c] = v;
    }
  } else if (kMode == RM_RESIDUAL) {
#pragma unroll
    for (int c = 0; c < 3; ++c) a.y[3 * (size_t)r + c] = a.b[3 * (size_t)r + c] - s[c];
  } else {
    // x_r = solve3x3(D_r, b_r - sum, 0)                         (Colored_GS_MF_Kernel, cublas.py:129-135)
    float rhs[3], xs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) rhs[c] = a.b[3 * (size_t)r + c] - s[c];
    const int bad = solve3x3(D, rhs, xs);
    if (bad && a.not_spd != nullptr) atomicAdd(a.not_spd, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) a.y[3 * (size_t)r + c] = xs[c];
  }
}

static RowArgs base_args(const FembLevel* l, const float* values) {
  RowArgs a;
  a.L_ptr = l->L_ptr; a.L_col = l->L_col; a.U_ptr = l->U_ptr; a.U_col = l->U_col;
  a.values = values; a.off_l = l->off_l; a.off_d = l->off_d; a.off_u = l->off_u;
  a.row_begin = 0; a.row_end = l->n_vert; a.parts = FEMB_PART_ALL;
  a.x = nullptr; a.b = nullptr; a.y = nullptr; a.alpha = 1.f; a.beta = 0.f; a.not_spd = nullptr; a.active = nullptr;
  return a;
}

template <int kMode>
static int launch_rows(const RowArgs& a, cudaStream_t st) {
  const int n = a.row_end - a.row_begin;
  if (n <= 0) return FEMB_OK;
  row_pass_kernel<kMode><<<cdiv(n, RP_ROWS), RP_THREADS, 0, st>>>(a);
  count_launch();
  return check_launch("row_pass_kernel");
}

// ---- vector kernels -----------------------------------------------------------------------------
__global__ void axpby_kernel(int n, float a, const float* __restrict__ x, float b, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b != 0.f ? b * y[i] : 0.f) + (a != 0.f ? a * x[i] : 0.f);
}
__global__ void copy_kernel(int n, const float* __restrict__ x, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i];
}
__global__ void fill_kernel(int n, float v, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = v;
}
__global__ void update_x_kernel(int n, const int* __restrict__ i2v, const float* __restrict__ dx, float* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = __ldg(i2v + i);
#pragma unroll
  for (int c = 0; c < 3; ++c) x[3 * (size_t)v + c] = x[3 * (size_t)v + c] + dx[3 * (size_t)i + c];
}
__global__ void permute_kernel(int n, const int* __restrict__ map, const float* __restrict__ src, float* __restrict__ dst) {
  // dst[map[i]] = src[i]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = __ldg(map + i);
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[3 * (size_t)j + c] = src[3 * (size_t)i + c];
}
__global__ void inf_norm_kernel(int n, const float* __restrict__ x, float* __restrict__ out) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
__global__ void __launch_bounds__(256) dot_partial_kernel(int n, const float* __restrict__ x, const float* __restrict__ y,
                                                         float* __restrict__ partial, const float* active) {
  __shared__ float red[32];
  if (active != nullptr && *active == 0.f) return;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += x[i] * y[i];
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) dot_final_kernel(int nb, const float* __restrict__ partial, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = s;
}

static int reduce_blocks(int n) {
  int b = (int)cdiv(n, 256 * 8);
  return b < 1 ? 1 : (b > 592 ? 592 : b);   // 4 CTAs x 148 SMs
}

// ---- CG scalar state machine (replaces the host logic of object.py:944-973) ---------------------
// state: 0 r1, 1 r0, 2 r (initial), 3 dot, 4 alpha, 5 beta, 6 active, 7 k, 8 iterations, 9 tol, 16.. partials
constexpr int CG_PART = 16;
__global__ void __launch_bounds__(256) cg_scalar_kernel(int stage, int nb, float* st) {
  __shared__ float red[32];
  float acc = 0.f;
  const bool active = st[6] != 0.f || stage == 0;
  if (active)
    for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += st[CG_PART + i];
  const float s = block_sum(acc, red);
  if (threadIdx.x != 0 || !active) return;
  if (stage == 0) {                 // r1 = |b|^2, object.py:940-947
    st[0] = s; st[1] = 0.f; st[2] = s; st[7] = 1.f;
    bool act = !(s < 1e-7f);
    act = act && (s > st[9] * s) && (1.f <= st[8]);      // while (r1 > tol*r and k <= iterations)
    st[6] = act ? 1.f : 0.f;
    st[5] = 0.f;
  } else if (stage == 1) {          // dot = p.Ap, object.py:959-965
    st[3] = s;
    if (s < 1e-10f) { st[6] = 0.f; return; }
    st[4] = st[0] / s;
  } else {                          // r1 = |r|^2, object.py:968-973 and the next loop test
    st[1] = st[0]; st[0] = s; st[7] += 1.f;
    const bool act = (s > st[9] * st[2]) && (st[7] <= st[8]);
    st[6] = act ? 1.f : 0.f;
    st[5] = s / st[1];
  }
}
__global__ void cg_update_p_kernel(int n, const float* __restrict__ r, float* __restrict__ p, const float* st) {
  if (st[6] == 0.f) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = st[7] > 1.f ? p[i] * st[5] + r[i] : r[i];              // scal + axpy / copy, object.py:949-954
}
__global__ void cg_update_xr_kernel(int n, const float* __restrict__ p, const float* __restrict__ ap, float* __restrict__ x,
                                    float* __restrict__ r, const float* st) {
  if (st[6] == 0.f) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float alpha = st[4];
  x[i] = x[i] + alpha * p[i];                                    // object.py:966-967
  r[i] = r[i] + (-alpha) * ap[i];
}

// ---- grid transfer: y = alpha * T x + beta * y with scalar weights ------------------------------
__global__ void __launch_bounds__(256) transfer_kernel(int nrow, const int* __restrict__ off, const int* __restrict__ col,
                                                      const float* __restrict__ w, float alpha, const float* __restrict__ x,
                                                      float beta, float* __restrict__ y) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  float s[3] = {0.f, 0.f, 0.f};
  for (int p = __ldg(off + r); p < __ldg(off + r + 1); ++p) {
    const int c = __ldg(col + p);
    const float ww = __ldg(w + p);
    s[0] += ww * x[3 * (size_t)c]; s[1] += ww * x[3 * (size_t)c + 1]; s[2] += ww * x[3 * (size_t)c + 2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = alpha * s[c];
    if (beta != 0.f) v += beta * y[3 * (size_t)r + c];
    y[3 * (size_t)r + c] = v;
  }
}

// ---- Galerkin product on fixed patterns -----------------------------------------------------------
constexpr int GK_THREADS = 64;
constexpr int GK_SLOTS = 27;

__device__ __forceinline__ int find_col(const int* cols, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cols + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// T = A_f * Us_hat : thread per fine row, private accumulators acc[slot*9+k][thread] in shared memory
__global__ void __launch_bounds__(GK_THREADS) galerkin_ap_kernel(int n_fine, const int* __restrict__ L_ptr,
                                                                const int* __restrict__ L_col, const int* __restrict__ U_ptr,
                                                                const int* __restrict__ U_col, const float* __restrict__ values,
                                                                int off_l, int off_d, int off_u, const int* __restrict__ p_off,
                                                                const int* __restrict__ p_col, const float* __restrict__ p_w,
                                                                const int* __restrict__ t_ptr, const int* __restrict__ t_col,
                                                                float* __restrict__ t_val) {
  extern __shared__ float s_acc[];
  const int i = blockIdx.x * GK_THREADS + threadIdx.x;
  if (i >= n_fine) return;
  float* acc = s_acc + threadIdx.x;
  const int t0 = __ldg(t_ptr + i), tn = __ldg(t_ptr + i + 1) - t0;
  for (int s = 0; s < tn * 9; ++s) acc[s * GK_THREADS] = 0.f;
  const int l0 = __ldg(L_ptr + i), l1 = __ldg(L_ptr + i + 1), u0 = __ldg(U_ptr + i), u1 = __ldg(U_ptr + i + 1);
  const int total = (l1 - l0) + 1 + (u1 - u0);
  for (int jj = 0; jj < total; ++jj) {          // blocks of row i in column order: L, D, U
    int j; size_t g;
    if (jj < l1 - l0) { j = __ldg(L_col + l0 + jj); g = (size_t)off_l + l0 + jj; }
    else if (jj == l1 - l0) { j = i; g = (size_t)off_d + i; }
    else { const int u = u0 + (jj - (l1 - l0) - 1); j = __ldg(U_col + u); g = (size_t)off_u + u; }
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = __ldg(values + g * 9 + k);
    for (int p = __ldg(p_off + j); p < __ldg(p_off + j + 1); ++p) {
      const int J = __ldg(p_col + p);
      const float w = __ldg(p_w + p);
      const int slot = find_col(t_col + t0, tn, J);
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[(slot * 9 + k) * GK_THREADS] += m[k] * w;
    }
  }
  for (int s = 0; s < tn * 9; ++s) t_val[(size_t)t0 * 9 + s] = acc[s * GK_THREADS];
}

// A_c = Ut_hat * T (+ spd_matrix33f), written to the [L;D;U] slot of each (row,col)-sorted block
__global__ void __launch_bounds__(GK_THREADS) galerkin_rt_kernel(int n_coarse, const int* __restrict__ r_off,
                                                                const int* __restrict__ r_col, const float* __restrict__ r_w,
                                                                const int* __restrict__ t_ptr, const int* __restrict__ t_col,
                                                                const float* __restrict__ t_val, const int* __restrict__ c_ptr,
                                                                const int* __restrict__ c_col, const int* __restrict__ c_slot,
                                                                int spd, float spd_value, float* __restrict__ out) {
  extern __shared__ float s_acc[];
  const int I = blockIdx.x * GK_THREADS + threadIdx.x;
  if (I >= n_coarse) return;
  float* acc = s_acc + threadIdx.x;
  const int c0 = __ldg(c_ptr + I), cn = __ldg(c_ptr + I + 1) - c0;
  for (int s = 0; s < cn * 9; ++s) acc[s * GK_THREADS] = 0.f;
  for (int p = __ldg(r_off + I); p < __ldg(r_off + I + 1); ++p) {
    const int i = __ldg(r_col + p);
    const float w = __ldg(r_w + p);
    for (int t = __ldg(t_ptr + i); t < __ldg(t_ptr + i + 1); ++t) {
      const int slot = find_col(c_col + c0, cn, __ldg(t_col + t));
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[(slot * 9 + k) * GK_THREADS] += w * __ldg(t_val + (size_t)t * 9 + k);
    }
  }
  for (int s = 0; s < cn; ++s) {
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = acc[(s * 9 + k) * GK_THREADS];
    if (spd) spd_block(m, spd_value);
    float* d = out + (size_t)__ldg(c_slot + c0 + s) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) d[k] = m[k];
  }
}

__global__ void __launch_bounds__(256) spd_kernel(int n, float value, float* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = vals[(size_t)i * 9 + k];
  spd_block(m, value);
#pragma unroll
  for (int k = 0; k < 9; ++k) vals[(size_t)i * 9 + k] = m[k];
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_spmv(const FembLevel* lvl, const float* values, int parts, float alpha, const float* x, float beta, float* y,
              void* stream) {
  FEMB_REQUIRE(lvl && values && x && y, "null pointer");
  FEMB_REQUIRE(parts > 0 && parts <= FEMB_PART_ALL, "parts");
  FEMB_REQUIRE(x != y, "x and y must not alias");
  RowArgs a = base_args(lvl, values);
  a.parts = parts; a.alpha = alpha; a.beta = beta; a.x = x; a.y = y;
  return launch_rows<RM_SPMV>(a, as_stream(stream));
}

int femb_residual(const FembLevel* lvl, const float* values, const float* b, const float* x, float* r, void* stream) {
  FEMB_REQUIRE(lvl && values && b && x && r, "null pointer");
  FEMB_REQUIRE(x != r, "x and r must not alias");
  RowArgs a = base_args(lvl, values);
  a.x = x; a.b = b; a.y = r;
  return launch_rows<RM_RESIDUAL>(a, as_stream(stream));
}

int femb_gauss_seidel(const FembLevel* lvl, const float* values, const float* b, float* dx, int iterations, int* not_spd,
                      void* stream) {
  FEMB_REQUIRE(lvl && values && b && dx, "null pointer");
  FEMB_REQUIRE(iterations >= 0, "iterations");
  cudaStream_t st = as_stream(stream);
  if (lvl->n_vert == 0) return FEMB_OK;
  cudaMemsetAsync(dx, 0, sizeof(float) * 3 * (size_t)lvl->n_vert, st);           // object.py:887
  RowArgs a = base_args(lvl, values);
  a.parts = FEMB_PART_L | FEMB_PART_U; a.x = dx; a.y = dx; a.b = b; a.not_spd = not_spd;
  for (int it = 0; it < iterations; ++it) {
    for (int c = 7; c >= 0; --c) {                                                // backward sweep, object.py:893-898
      a.row_begin = lvl->color_start[c]; a.row_end = lvl->color_start[c + 1];
      // first half-sweep of the first iteration: dx == 0, rows of colour c only see higher colours
      a.parts = (it == 0) ? FEMB_PART_U : (FEMB_PART_L | FEMB_PART_U);
      int rc = launch_rows<RM_GS>(a, st);
      if (rc) return rc;
    }
    a.parts = FEMB_PART_L | FEMB_PART_U;
    for (int c = 0; c < 8; ++c) {                                                 // forward sweep, object.py:903-908
      a.row_begin = lvl->color_start[c]; a.row_end = lvl->color_start[c + 1];
      int rc = launch_rows<RM_GS>(a, st);
      if (rc) return rc;
    }
  }
  return FEMB_OK;
}

int femb_jacobi(const FembLevel* lvl, const float* values, const float* b, float* dx, float* scratch, int iterations,
                int* not_spd, void* stream) {
  FEMB_REQUIRE(lvl && values && b && dx && scratch, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (lvl->n_vert == 0) return FEMB_OK;
  const size_t bytes = sizeof(float) * 3 * (size_t)lvl->n_vert;
  cudaMemsetAsync(dx, 0, bytes, st);                                              // object.py:870
  RowArgs a = base_args(lvl, values);
  a.b = b; a.not_spd = not_spd; a.x = dx; a.y = scratch;
  for (int it = 0; it < iterations; ++it) {
    a.parts = (it == 0) ? 0 : (FEMB_PART_L | FEMB_PART_U);                        // dx == 0 on the first pass
    int rc = launch_rows<RM_JACOBI>(a, st);
    if (rc) return rc;
    cudaMemcpyAsync(dx, scratch, bytes, cudaMemcpyDeviceToDevice, st);
  }
  return FEMB_OK;
}

int femb_cg(const FembLevel* lvl, const float* values, const float* b, float* dx, float* r, float* p, float* ap,
            float* state, int iterations, float tol, int with_init_x, void* stream) {
  FEMB_REQUIRE(lvl && values && b && dx && r && p && ap && state, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int n3 = 3 * lvl->n_vert;
  if (n3 == 0) return FEMB_OK;
  const int nb = reduce_blocks(n3);
  const unsigned vb = cdiv(n3, 256);
  fill_kernel<<<1, 64, 0, st>>>(16, 0.f, state);
  fill_kernel<<<1, 1, 0, st>>>(1, (float)iterations, state + 8);
  fill_kernel<<<1, 1, 0, st>>>(1, tol, state + 9);
  count_launch(3);
  RowArgs a = base_args(lvl, values);
  if (with_init_x) {                                                              // object.py:982-985
    a.x = dx; a.b = b; a.y = r;
    int rc = launch_rows<RM_RESIDUAL>(a, st);
    if (rc) return rc;
  } else {
    cudaMemsetAsync(dx, 0, sizeof(float) * n3, st);                              // object.py:932
    cudaMemcpyAsync(r, b, sizeof(float) * n3, cudaMemcpyDeviceToDevice, st);     // object.py:939
  }
  dot_partial_kernel<<<nb, 256, 0, st>>>(n3, r, r, state + CG_PART, nullptr);
  cg_scalar_kernel<<<1, 256, 0, st>>>(0, nb, state);
  count_launch(2);
  for (int it = 0; it < iterations; ++it) {
    cg_update_p_kernel<<<vb, 256, 0, st>>>(n3, r, p, state);
    count_launch();
    a = base_args(lvl, values);
    a.x = p; a.y = ap; a.alpha = 1.f; a.beta = 0.f; a.active = state + 6;         // object.py:955-958
    int rc = launch_rows<RM_SPMV>(a, st);
    if (rc) return rc;
    dot_partial_kernel<<<nb, 256, 0, st>>>(n3, p, ap, state + CG_PART, state + 6);
    cg_scalar_kernel<<<1, 256, 0, st>>>(1, nb, state);
    cg_update_xr_kernel<<<vb, 256, 0, st>>>(n3, p, ap, dx, r, state);
    dot_partial_kernel<<<nb, 256, 0, st>>>(n3, r, r, state + CG_PART, state + 6);
    cg_scalar_kernel<<<1, 256, 0, st>>>(2, nb, state);
    count_launch(5);
  }
  return check_launch("cg");
}

int femb_spd_blocks(float* values, int n_blocks, float spd_value, void* stream) {
  FEMB_REQUIRE(values && n_blocks >= 0, "null pointer or negative size");
  if (n_blocks == 0) return FEMB_OK;
  spd_kernel<<<cdiv(n_blocks, 256), 256, 0, as_stream(stream)>>>(n_blocks, spd_value, values);
  count_launch();
  return check_launch("spd_kernel");
}

int femb_transfer_mv(const FembTransfer* t, float alpha, const float* x, float beta, float* y, void* stream) {
  FEMB_REQUIRE(t && x && y && t->offsets && t->columns && t->weights, "null pointer");
  if (t->nrow == 0) return FEMB_OK;
  transfer_kernel<<<cdiv(t->nrow, 256), 256, 0, as_stream(stream)>>>(t->nrow, t->offsets, t->columns, t->weights, alpha, x,
                                                                     beta, y);
  count_launch();
  return check_launch("transfer_kernel");
}

int femb_galerkin(const FembLevel* fine, const float* fine_values, const FembTransfer* ut_hat, const FembTransfer* us_hat,
                  const FembGalerkinPlan* plan, float* t_values, int spd, float spd_value, float* coarse_values,
                  void* stream) {
  FEMB_REQUIRE(fine && fine_values && ut_hat && us_hat && plan && t_values && coarse_values, "null pointer");
  FEMB_REQUIRE(plan->n_fine == fine->n_vert && us_hat->nrow == fine->n_vert && ut_hat->nrow == plan->n_coarse,
               "inconsistent sizes");
  static bool configured = false;
  const size_t smem = sizeof(float) * GK_THREADS * GK_SLOTS * 9;
  if (!configured) {
    cudaFuncSetAttribute(galerkin_ap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(galerkin_rt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
  if (plan->n_fine > 0) {
    galerkin_ap_kernel<<<cdiv(plan->n_fine, GK_THREADS), GK_THREADS, smem, st>>>(
        plan->n_fine, fine->L_ptr, fine->L_col, fine->U_ptr, fine->U_col, fine_values, fine->off_l, fine->off_d, fine->off_u,
        us_hat->offsets, us_hat->columns, us_hat->weights, plan->t_ptr, plan->t_col, t_values);
    count_launch();
  }
  if (plan->n_coarse > 0) {
    galerkin_rt_kernel<<<cdiv(plan->n_coarse, GK_THREADS), GK_THREADS, smem, st>>>(
        plan->n_coarse, ut_hat->offsets, ut_hat->columns, ut_hat->weights, plan->t_ptr, plan->t_col, t_values, plan->c_ptr,
        plan->c_col, plan->c_slot, spd, spd_value, coarse_values);
    count_launch();
  }
  return check_launch("galerkin");
}

int femb_axpby(int n, float a, const float* x, float b, float* y, void* stream) {
  FEMB_REQUIRE(n >= 0 && x && y, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  axpby_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, a, x, b, y);
  count_launch();
  return check_launch("axpby_kernel");
}
int femb_copy(int n, const float* x, float* y, void* stream) {
  FEMB_REQUIRE(n >= 0 && x && y, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  copy_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, x, y);
  count_launch();
  return check_launch("copy_kernel");
}
int femb_fill(int n, float v, float* y, void* stream) {
  FEMB_REQUIRE(n >= 0 && y, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  fill_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, v, y);
  count_launch();
  return check_launch("fill_kernel");
}
int femb_update_x(const FembLevel* lvl, const float* dx, float* x, void* stream) {
  FEMB_REQUIRE(lvl && dx && x && lvl->i2v, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  update_x_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->i2v, dx, x);
  count_launch();
  return check_launch("update_x_kernel");
}
int femb_v2i(const FembLevel* lvl, const float* vert, float* ordered, void* stream) {
  FEMB_REQUIRE(lvl && vert && ordered && lvl->v2i, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  permute_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->v2i, vert, ordered);
  count_launch();
  return check_launch("permute_kernel");
}
int femb_i2v(const FembLevel* lvl, const float* ordered, float* vert, void* stream) {
  FEMB_REQUIRE(lvl && vert && ordered && lvl->i2v, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  permute_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->i2v, ordered, vert);
  count_launch();
  return check_launch("permute_kernel");
}
int femb_inf_norm(int n, const float* x, float* out, void* stream) {
  FEMB_REQUIRE(n >= 0 && x && out, "null pointer or negative size");
  cudaMemsetAsync(out, 0, sizeof(float), as_stream(stream));
  if (n == 0) return FEMB_OK;
  int b = (int)cdiv(n, 256 * 4);
  b = b > 1184 ? 1184 : b;
  inf_norm_kernel<<<b, 256, 0, as_stream(stream)>>>(n, x, out);
  count_launch();
  return check_launch("inf_norm_kernel");
}
int femb_dot(int n, const float* x, const float* y, float* scratch, float* out, void* stream) {
  FEMB_REQUIRE(n >= 0 && x && y && scratch && out, "null pointer or negative size");
  const int nb = reduce_blocks(n);
  dot_partial_kernel<<<nb, 256, 0, as_stream(stream)>>>(n, x, y, scratch, nullptr);
  dot_final_kernel<<<1, 256, 0, as_stream(stream)>>>(nb, scratch, out);
  count_launch(2);
  return check_launch("dot");
}

}  // extern "C"
