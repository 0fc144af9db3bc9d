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
// HBM into shared memory with bulk copies, then 4 or 8 threads per row walk the staged blocks
// (lanes = consecutive rows, so the x[col] gathers of a lattice mesh coalesce), and an epilogue turns
// the row sum into y / residual / a solve3x3 update.  A colour of the 8-colouring is a contiguous row
// range, so one symmetric Gauss-Seidel iteration is 16 such launches that update dx in place: rows of
// colour c read neighbours of lower colours (old values in the backward sweep, new in the forward
// sweep) and higher colours (the opposite) -- exactly b - L dx_old - U dx_new of object.py:890-908
// without the per-colour matrix rebuilds.  Matrix bytes move once per half-sweep.
#include "femb_common.cuh"

namespace femb {

constexpr int RP_ROWS = 32;                 // rows per CTA   } defaults of the templates;
constexpr int RP_SUBS = 4;                  // threads per row } the launchers pick 16 x 8 (FEMB_ROWS_DISPATCH)
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

// (mbarrier / cp.async.bulk helpers: femb_common.cuh)

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor is still running; griddep_wait() blocks until the predecessor has completed and its
// writes are visible, griddep_launch() lets the successor start.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int kSubs, int kRows = RP_ROWS>
struct RowSmemT {
  static constexpr int CAP = kRows * 27;     // staged off-diagonal blocks per CTA (26 neighbours + slack)
  alignas(128) float val[CAP * 9 + 16];      // [L chunk | U chunk], each starting on a 16-byte boundary
  alignas(16) int col[CAP + 16];
  alignas(16) float diag[kRows * 9 + 8];
  int ptr[2][kRows + 1];
  alignas(8) uint64_t bar;
  // the partial row sums of the kSubs threads of a row are exchanged through `val` once everybody is done reading it
  static_assert(sizeof(float) * kSubs * kRows * 3 <= sizeof(float) * (CAP * 9 + 16), "red does not fit");
};
using RowSmem = RowSmemT<RP_SUBS>;

// One block of <= RP_ROWS consecutive rows.  kCoherent: x is read through L2 (ld.global.cg) instead of the per-SM L1 (for a
// caller whose other CTAs update x during the same launch; the shipped kernels do not).
// The values / L_col / U_col arrays must be 16-byte aligned and readable 16 bytes past their end
// (the bulk copies are rounded out to 16-byte granules).
template <int kMode, bool kCoherent, bool kPdl = false, int kSubs = RP_SUBS, int kRows = RP_ROWS>
__device__ __forceinline__ void row_block(const RowArgs& a, int r0, int nrow, RowSmemT<kSubs, kRows>& sm, uint32_t& phase) {
  const int tid = threadIdx.x;
  const bool useL = (a.parts & FEMB_PART_L) != 0, useU = (a.parts & FEMB_PART_U) != 0;
  const bool needD = (kMode == RM_GS || kMode == RM_JACOBI) || (a.parts & FEMB_PART_D) != 0;
  if (tid <= nrow) {
    sm.ptr[0][tid] = useL ? __ldg(a.L_ptr + r0 + tid) : 0;
    sm.ptr[1][tid] = useU ? __ldg(a.U_ptr + r0 + tid) : 0;
  }
  __syncthreads();
  const int L0 = sm.ptr[0][0], U0 = sm.ptr[1][0];
  const int nL = sm.ptr[0][nrow] - L0, nU = sm.ptr[1][nrow] - U0;
  const bool staged = (nL + nU) <= (RowSmemT<kSubs, kRows>::CAP);
  // float / int offsets of the wanted data inside the staged (16-byte rounded) chunks
  const size_t gL = ((size_t)a.off_l + L0) * 9, gU = ((size_t)a.off_u + U0) * 9, gD = ((size_t)a.off_d + r0) * 9;
  const int preL = (int)(gL & 3), preU = (int)(gU & 3), preD = (int)(gD & 3);
  const int baseU = (preL + nL * 9 + 3) & ~3;
  const int cpreL = L0 & 3, cpreU = U0 & 3;
  const int cbaseU = (cpreL + nL + 3) & ~3;
  if (tid == 0) {
    uint32_t bytes = 0;
    const uint32_t bL = (uint32_t)((preL + nL * 9) * 4 + 15) & ~15u, bU = (uint32_t)((preU + nU * 9) * 4 + 15) & ~15u;
    const uint32_t bD = (uint32_t)((preD + nrow * 9) * 4 + 15) & ~15u;
    const uint32_t cL = (uint32_t)((cpreL + nL) * 4 + 15) & ~15u, cU = (uint32_t)((cpreU + nU) * 4 + 15) & ~15u;
    if (staged && nL > 0) bytes += bL + cL;
    if (staged && nU > 0) bytes += bU + cU;
    if (needD) bytes += bD;
    mbar_expect_tx(&sm.bar, bytes);
    if (staged && nL > 0) {
      bulk_g2s(sm.val, a.values + (gL - preL), bL, &sm.bar);
      bulk_g2s(sm.col, a.L_col + (L0 - cpreL), cL, &sm.bar);
    }
    if (staged && nU > 0) {
      bulk_g2s(sm.val + baseU, a.values + (gU - preU), bU, &sm.bar);
      bulk_g2s(sm.col + cbaseU, a.U_col + (U0 - cpreU), cU, &sm.bar);
    }
    if (needD) bulk_g2s(sm.diag, a.values + (gD - preD), bD, &sm.bar);
  }
  if (kPdl) {
    // Everything above reads only the matrix (constant while a sweep runs): with a programmatic launch it overlaps
    // the tail of the previous colour.  x / b / y are touched only after the predecessor has completed.
    griddep_wait();
    griddep_launch();
  }
  // the right-hand side of this thread's row (epilogue) travels while the staging is in flight
  float bv[3] = {0.f, 0.f, 0.f};
  if ((kMode == RM_GS || kMode == RM_JACOBI || kMode == RM_RESIDUAL) && tid < nrow) {
    const float* bp = a.b + 3 * (size_t)(r0 + tid);
    bv[0] = __ldg(bp); bv[1] = __ldg(bp + 1); bv[2] = __ldg(bp + 2);
  }
  mbar_wait(&sm.bar, phase);
  phase ^= 1u;

  const float* vL = sm.val + preL;
  const float* vU = sm.val + baseU + preU;
  const int* cLs = sm.col + cpreL;
  const int* cUs = sm.col + cbaseU + cpreU;
  const int rl = tid % kRows, sub = tid / kRows;
  float acc[3] = {0.f, 0.f, 0.f};
  if (rl < nrow) {
    const int lb = sm.ptr[0][rl] - L0, cntL = sm.ptr[0][rl + 1] - sm.ptr[0][rl];
    const int ub = sm.ptr[1][rl] - U0, cntU = sm.ptr[1][rl + 1] - sm.ptr[1][rl];
    const int n = cntL + cntU;
    if (staged) {
      constexpr int MAXB = (27 + kSubs - 1) / kSubs;     // blocks per thread: one trip
      for (int j0 = sub; j0 < n; j0 += MAXB * kSubs) {
        const float* m[MAXB];
        float xv[MAXB][3];
#pragma unroll
        for (int u = 0; u < MAXB; ++u) {                 // issue every x gather of this thread first
          const int j = j0 + u * kSubs;
          const bool ok = j < n;
          const int jj = ok ? j : j0;
          m[u] = (jj < cntL) ? vL + (lb + jj) * 9 : vU + (ub + jj - cntL) * 9;
          const int c = (jj < cntL) ? cLs[lb + jj] : cUs[ub + jj - cntL];
          const float* xp = a.x + 3 * (size_t)c;
          if (kCoherent) { xv[u][0] = __ldcg(xp); xv[u][1] = __ldcg(xp + 1); xv[u][2] = __ldcg(xp + 2); }
          else { xv[u][0] = xp[0]; xv[u][1] = xp[1]; xv[u][2] = xp[2]; }
          if (!ok) { xv[u][0] = 0.f; xv[u][1] = 0.f; xv[u][2] = 0.f; }
        }
#pragma unroll
        for (int u = 0; u < MAXB; ++u) {
          acc[0] += m[u][0] * xv[u][0] + m[u][1] * xv[u][1] + m[u][2] * xv[u][2];
          acc[1] += m[u][3] * xv[u][0] + m[u][4] * xv[u][1] + m[u][5] * xv[u][2];
          acc[2] += m[u][6] * xv[u][0] + m[u][7] * xv[u][1] + m[u][8] * xv[u][2];
        }
      }
    } else {
      for (int j = sub; j < n; j += kSubs) {           // oversized rows: straight from global memory
        const bool isL = j < cntL;
        const size_t g = isL ? ((size_t)a.off_l + L0 + lb + j) : ((size_t)a.off_u + U0 + ub + (j - cntL));
        const int col = isL ? __ldg(a.L_col + L0 + lb + j) : __ldg(a.U_col + U0 + ub + (j - cntL));
        float m[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) m[k] = __ldg(a.values + g * 9 + k);
        const float x0 = __ldcg(a.x + 3 * (size_t)col), x1 = __ldcg(a.x + 3 * (size_t)col + 1), x2 = __ldcg(a.x + 3 * (size_t)col + 2);
        acc[0] += m[0] * x0 + m[1] * x1 + m[2] * x2;
        acc[1] += m[3] * x0 + m[4] * x1 + m[5] * x2;
        acc[2] += m[6] * x0 + m[7] * x1 + m[8] * x2;
      }
    }
  }
  __syncthreads();                                         // all reads of the staged blocks are done
  float (*red)[kRows][3] = reinterpret_cast<float (*)[kRows][3]>(sm.val);
  red[sub][rl][0] = acc[0]; red[sub][rl][1] = acc[1]; red[sub][rl][2] = acc[2];
  __syncthreads();
  if (tid < nrow) {
    const int r = r0 + tid;
    float s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s[c] = (red[0][tid][c] + red[1][tid][c]) + (red[2][tid][c] + red[3][tid][c]);
      if (kSubs == 8) s[c] += (red[4][tid][c] + red[5][tid][c]) + (red[6][tid][c] + red[7][tid][c]);
    }
    const float* D = sm.diag + preD + tid * 9;
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
      for (int c = 0; c < 3; ++c) a.y[3 * (size_t)r + c] = bv[c] - s[c];
    } else {
      // x_r = solve3x3(D_r, b_r - sum, 0)                       (Colored_GS_MF_Kernel, cublas.py:129-135)
      float rhs[3], xs[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c) rhs[c] = bv[c] - s[c];
      const int bad = solve3x3(D, rhs, xs);
      if (bad && a.not_spd != nullptr) atomicAdd(a.not_spd, 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) a.y[3 * (size_t)r + c] = xs[c];
    }
  }
}

template <int kMode, int kSubs, int kRows>
__global__ void __launch_bounds__(kRows * kSubs) row_pass_kernel(RowArgs a) {
  if (a.active != nullptr && *a.active == 0.f) return;
  __shared__ RowSmemT<kSubs, kRows> sm;
  if (threadIdx.x == 0) mbar_init(&sm.bar, 1);
  uint32_t phase = 0;
  const int r0 = a.row_begin + blockIdx.x * kRows;
  row_block<kMode, false, false, kSubs, kRows>(a, r0, min(kRows, a.row_end - r0), sm, phase);
}

// One colour of a Gauss-Seidel sweep, launchable as a programmatic dependent of the previous colour: the matrix
// staging (pointer reads + bulk copies) runs before griddepcontrol.wait, i.e. during the previous colour's tail.
// kRows rows per CTA x kSubs threads per row.  Measured per colour launch at 1 M hexahedra: 32x4 32.8 us, 32x8 34.7 us
// (same CTAs per SM, twice the warps: no gain), 16x8 28.7 us (default), 16x4 29.1, 8x4 28.4, 8x8 29.2, and 42.6 us for a
// persistent double-buffered variant with 3 CTAs per SM (removed) -- what bounds the pass is the number of independent
// row blocks in flight per SM (each one a chain of pointer load -> staging -> L2 gathers of x -> 3x3 solve), so smaller
// blocks win until the per-block overheads take over.
template <int kSubs, int kRows>
__global__ void __launch_bounds__(kRows * kSubs) gs_color_kernel(RowArgs a) {
  __shared__ RowSmemT<kSubs, kRows> sm;
  if (threadIdx.x == 0) mbar_init(&sm.bar, 1);
  uint32_t phase = 0;
  const int r0 = a.row_begin + blockIdx.x * kRows;
  row_block<RM_GS, false, true, kSubs, kRows>(a, r0, min(kRows, a.row_end - r0), sm, phase);
}

static RowArgs base_args(const FembLevel* l, const float* values) {
  RowArgs a;
  a.L_ptr = l->L_ptr; a.L_col = l->L_col; a.U_ptr = l->U_ptr; a.U_col = l->U_col;
  a.values = values; a.off_l = l->off_l; a.off_d = l->off_d; a.off_u = l->off_u;
  a.row_begin = 0; a.row_end = l->n_vert; a.parts = FEMB_PART_ALL;
  a.x = nullptr; a.b = nullptr; a.y = nullptr; a.alpha = 1.f; a.beta = 0.f; a.not_spd = nullptr; a.active = nullptr;
  return a;
}

// rows per CTA x threads per row of the one-block-per-CTA kernels (measured at 1 M hexahedra, per Gauss-Seidel colour
// launch: 32x4 32.8 us, 32x8 34.7, 16x4 29.3, 16x8 28.8 -- smaller blocks mean more independent chains in flight per SM)
static int g_gs_rows = 16;
static int g_gs_subs = 8;
#define FEMB_ROWS_DISPATCH(KERNEL_EXPR)                                        \
  do {                                                                         \
    if (g_gs_rows == 8 && g_gs_subs == 8) { constexpr int R = 8, S = 8; KERNEL_EXPR; }        \
    else if (g_gs_rows == 8) { constexpr int R = 8, S = 4; KERNEL_EXPR; }                      \
    else if (g_gs_rows == 16 && g_gs_subs == 8) { constexpr int R = 16, S = 8; KERNEL_EXPR; }  \
    else if (g_gs_rows == 16) { constexpr int R = 16, S = 4; KERNEL_EXPR; }                    \
    else if (g_gs_subs == 8) { constexpr int R = 32, S = 8; KERNEL_EXPR; }                     \
    else { constexpr int R = 32, S = 4; KERNEL_EXPR; }                                         \
  } while (0)

template <int kMode>
static int launch_rows(const RowArgs& a, cudaStream_t st) {
  const int n = a.row_end - a.row_begin;
  if (n <= 0) return FEMB_OK;
  FEMB_ROWS_DISPATCH((row_pass_kernel<kMode, S, R><<<cdiv(n, R), R * S, 0, st>>>(a)));
  count_launch();
  return check_launch("row_pass_kernel");
}

// pdl: the launch may overlap the tail of the previous kernel in the stream (which must not write the matrix)
static int g_no_pdl = 0;
static int launch_gs_color(const RowArgs& a, bool pdl, cudaStream_t st) {
  const int n = a.row_end - a.row_begin;
  if (n <= 0) return FEMB_OK;
  cudaLaunchConfig_t cfg = {};
  const int rows = g_gs_rows;
  cfg.gridDim = dim3(cdiv(n, rows)); cfg.blockDim = dim3(rows * g_gs_subs); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (pdl && !g_no_pdl) ? 1 : 0;
  cudaError_t e = cudaSuccess;
  FEMB_ROWS_DISPATCH((e = cudaLaunchKernelEx(&cfg, gs_color_kernel<S, R>, a)));
  count_launch();
  if (e != cudaSuccess) {
    set_error("gs_color_kernel: %s", cudaGetErrorString(e));
    return FEMB_E_CUDA;
  }
  return check_launch("gs_color_kernel");
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
// updateM + updateV + scal_ x2 + updateX of the reference's Adam loop (cublas.py:172-194, object.py:1090-1094) in one pass
__global__ void __launch_bounds__(256) adam_kernel(int n, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   float* __restrict__ m_hat, float* __restrict__ v_hat, float* __restrict__ x,
                                                   float a1, float b1, float a2, float b2, float s1, float s2, float lr, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = a1 * gi + b1 * m[i];
  const float vi = a2 * (gi * gi) + b2 * v[i];
  const float mh = mi * s1, vh = vi * s2;
  m[i] = mi; v[i] = vi; m_hat[i] = mh; v_hat[i] = vh;
  x[i] -= lr * (mh / (sqrtf(vh) + eps));
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
// A_c = Ut_hat (A_f Us_hat) in two streaming stages with precomputed slot plans (no searches, no atomics,
// fixed summation order = the row-major order of _bsr_mm_compute_values, sparse.py:873-901):
//   stage 1  T = A_f Us_hat : 32 fine rows per CTA, the rows' blocks staged by bulk copies like the row
//            pass, 9 threads per row (one per block entry) accumulate the <= 27 T blocks of the row
//   stage 2  A_c = Ut_hat T : 32 coarse vertices per CTA (vertex order, so the <= 8 parents of a fine row
//            are processed close in time and T is read from HBM once), 9 threads per coarse row,
//            spd_matrix33f in the epilogue, output written to the [L;D;U] slot of every block.
constexpr int GA_ROWS = 16;                 // 16 rows = 37 KB of staging + accumulators per CTA: 6 CTAs per SM (32 rows: 2 CTAs)
constexpr int GA_THREADS = GA_ROWS * 9;
constexpr int GA_SLOTS = 27;
constexpr int GA_CAP = GA_ROWS * 27;        // staged off-diagonal blocks per CTA

struct GaSmem {
  alignas(128) float val[GA_CAP * 9 + 16];
  alignas(16) int col[GA_CAP + 16];
  alignas(16) float diag[GA_ROWS * 9 + 8];
  alignas(16) unsigned char tsl[(GA_CAP + GA_ROWS) * 8 + 64];   // t_slot of the L chunk | U chunk | D chunk
  int ptr[2][GA_ROWS + 1];
  float acc[GA_ROWS * GA_SLOTS * 9];
  alignas(8) uint64_t bar;
};

template <bool kDyadic>
__global__ void __launch_bounds__(GA_THREADS) galerkin_ap_kernel(RowArgs a, int n_fine, const int* __restrict__ p_off,
                                                                const float* __restrict__ p_w,
                                                                const unsigned char* __restrict__ t_slot,
                                                                const int* __restrict__ t_ptr, float* __restrict__ t_val) {
  extern __shared__ __align__(128) unsigned char ga_raw[];
  GaSmem& sm = *reinterpret_cast<GaSmem*>(ga_raw);
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * GA_ROWS;
  const int nrow = min(GA_ROWS, n_fine - r0);
  if (tid == 0) mbar_init(&sm.bar, 1);
  if (tid <= nrow) {
    sm.ptr[0][tid] = __ldg(a.L_ptr + r0 + tid);
    sm.ptr[1][tid] = __ldg(a.U_ptr + r0 + tid);
  }
  for (int i = tid; i < GA_ROWS * GA_SLOTS * 9; i += GA_THREADS) sm.acc[i] = 0.f;
  __syncthreads();
  const int L0 = sm.ptr[0][0], U0 = sm.ptr[1][0];
  const int nL = sm.ptr[0][nrow] - L0, nU = sm.ptr[1][nrow] - U0;
  const bool staged = (nL + nU) <= GA_CAP;
  const size_t sL = (size_t)a.off_l + L0, sU = (size_t)a.off_u + U0, sD = (size_t)a.off_d + r0;   // first [L;D;U] slots
  const size_t gL = sL * 9, gU = sU * 9, gD = sD * 9;
  const int preL = (int)(gL & 3), preU = (int)(gU & 3), preD = (int)(gD & 3);
  const int baseU = (preL + nL * 9 + 3) & ~3;
  const int cpreL = L0 & 3, cpreU = U0 & 3;
  const int cbaseU = (cpreL + nL + 3) & ~3;
  // t_slot chunks (8 bytes per block): byte offsets inside sm.tsl
  const int tpreL = (int)((sL * 8) & 15), tpreU = (int)((sU * 8) & 15), tpreD = (int)((sD * 8) & 15);
  const int tbaseU = (tpreL + nL * 8 + 15) & ~15;
  const int tbaseD = tbaseU + ((tpreU + nU * 8 + 15) & ~15);
  if (tid == 0) {
    const uint32_t bL = (uint32_t)((preL + nL * 9) * 4 + 15) & ~15u, bU = (uint32_t)((preU + nU * 9) * 4 + 15) & ~15u;
    const uint32_t bD = (uint32_t)((preD + nrow * 9) * 4 + 15) & ~15u;
    const uint32_t cL = (uint32_t)((cpreL + nL) * 4 + 15) & ~15u, cU = (uint32_t)((cpreU + nU) * 4 + 15) & ~15u;
    const uint32_t tL = (uint32_t)(tpreL + nL * 8 + 15) & ~15u, tU = (uint32_t)(tpreU + nU * 8 + 15) & ~15u;
    const uint32_t tD = (uint32_t)(tpreD + nrow * 8 + 15) & ~15u;
    uint32_t bytes = bD + tD;
    if (staged && nL > 0) bytes += bL + cL + tL;
    if (staged && nU > 0) bytes += bU + cU + tU;
    mbar_expect_tx(&sm.bar, bytes);
    if (staged && nL > 0) {
      bulk_g2s(sm.val, a.values + (gL - preL), bL, &sm.bar);
      bulk_g2s(sm.col, a.L_col + (L0 - cpreL), cL, &sm.bar);
      bulk_g2s(sm.tsl, t_slot + (sL * 8 - tpreL), tL, &sm.bar);
    }
    if (staged && nU > 0) {
      bulk_g2s(sm.val + baseU, a.values + (gU - preU), bU, &sm.bar);
      bulk_g2s(sm.col + cbaseU, a.U_col + (U0 - cpreU), cU, &sm.bar);
      bulk_g2s(sm.tsl + tbaseU, t_slot + (sU * 8 - tpreU), tU, &sm.bar);
    }
    bulk_g2s(sm.diag, a.values + (gD - preD), bD, &sm.bar);
    bulk_g2s(sm.tsl + tbaseD, t_slot + (sD * 8 - tpreD), tD, &sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  const int rl = tid / 9, k = tid - rl * 9;
  if (rl >= nrow) return;
  const int i = r0 + rl;
  const int lb = sm.ptr[0][rl] - L0, cntL = sm.ptr[0][rl + 1] - sm.ptr[0][rl];
  const int ub = sm.ptr[1][rl] - U0, cntU = sm.ptr[1][rl + 1] - sm.ptr[1][rl];
  float* acc = sm.acc + tid;                      // acc[slot][tid]: bank == lane, conflict-free
  // one fine block times the <= 8 parents of its column: acc[slot of parent in the T row] += weight * v.  The 8 plan bytes
  // (valid ones first, 255 = none) arrive in one 8-byte shared load.
  auto apply_dyadic = [&](float v, const unsigned char* ts) {
    const uint2 pk = *reinterpret_cast<const uint2*>(ts);
    unsigned w = pk.x;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      if (p == 4) w = pk.y;
      const unsigned b8 = w & 255u;
      w >>= 8;
      if (b8 == 255u) break;
      // weights 2^-e (e = 0..3) travel in bits 5-6 of the plan byte
      float* d = acc + (b8 & 31u) * GA_THREADS;
      *d = fmaf(v, __int_as_float(0x3f800000 - (int)((b8 & 0x60u) << 18)), *d);
    }
  };
  if (kDyadic && staged) {
    // fast path: everything is in shared memory; L blocks, D, U blocks = column order (sparse.py:873-901)
    const float* vL = sm.val + preL + lb * 9 + k;
    const unsigned char* tL = sm.tsl + tpreL + lb * 8;
    for (int jj = 0; jj < cntL; ++jj) apply_dyadic(vL[jj * 9], tL + jj * 8);
    apply_dyadic(sm.diag[preD + rl * 9 + k], sm.tsl + tbaseD + tpreD + rl * 8);
    const float* vU = sm.val + baseU + preU + ub * 9 + k;
    const unsigned char* tU = sm.tsl + tbaseU + tpreU + ub * 8;
    for (int jj = 0; jj < cntU; ++jj) apply_dyadic(vU[jj * 9], tU + jj * 8);
  } else {
    const int total = cntL + 1 + cntU;
    for (int jj = 0; jj < total; ++jj) {                       // column order: L blocks, D, U blocks
      float v;
      int j;
      const unsigned char* ts;
      if (jj < cntL) {
        const int loc = lb + jj;
        if (staged) { v = sm.val[preL + loc * 9 + k]; j = sm.col[cpreL + loc]; ts = sm.tsl + tpreL + loc * 8; }
        else { v = __ldg(a.values + (sL + loc) * 9 + k); j = __ldg(a.L_col + L0 + loc); ts = t_slot + (sL + loc) * 8; }
      } else if (jj == cntL) {
        v = sm.diag[preD + rl * 9 + k]; j = i; ts = sm.tsl + tbaseD + tpreD + rl * 8;
      } else {
        const int loc = ub + (jj - cntL - 1);
        if (staged) { v = sm.val[baseU + preU + loc * 9 + k]; j = sm.col[cbaseU + cpreU + loc]; ts = sm.tsl + tbaseU + tpreU + loc * 8; }
        else { v = __ldg(a.values + (sU + loc) * 9 + k); j = __ldg(a.U_col + U0 + loc); ts = t_slot + (sU + loc) * 8; }
      }
      if (kDyadic) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const unsigned b8 = ts[p];
          if (b8 == 255u) break;
          acc[(int)(b8 & 31u) * GA_THREADS] += v * __int_as_float((127 - (int)(b8 >> 5)) << 23);
        }
      } else {
        const int p0 = __ldg(p_off + j), np = __ldg(p_off + j + 1) - p0;
        for (int p = 0; p < np; ++p) acc[(int)(ts[p] & 31u) * GA_THREADS] += v * __ldg(p_w + p0 + p);
      }
    }
  }
  const int t0 = __ldg(t_ptr + i), tn = __ldg(t_ptr + i + 1) - t0;
  for (int sl = 0; sl < tn; ++sl) t_val[((size_t)t0 + sl) * 9 + k] = acc[sl * GA_THREADS];
}

__global__ void __launch_bounds__(GA_THREADS) galerkin_rt_kernel(int n_coarse, const int* __restrict__ v2i_c,
                                                                const int* __restrict__ r_off, const int* __restrict__ r_col,
                                                                const float* __restrict__ r_w,
                                                                const unsigned char* __restrict__ ut_q,
                                                                const int* __restrict__ t_ptr, const float* __restrict__ t_val,
                                                                const unsigned char* __restrict__ c_slot8,
                                                                const int* __restrict__ c_ptr, const int* __restrict__ c_slot,
                                                                int spd, float spd_value, float* __restrict__ out) {
  __shared__ float s_acc[GA_ROWS * GA_SLOTS * 9];
  const int tid = threadIdx.x;
  for (int i = tid; i < GA_ROWS * GA_SLOTS * 9; i += GA_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  const int rl = tid / 9, k = tid - rl * 9;
  const int v = blockIdx.x * GA_ROWS + rl;
  const bool live = v < n_coarse;
  int I = 0;
  float* acc = s_acc + tid;                       // acc[slot][tid]
  if (live) {
    I = __ldg(v2i_c + v);
    // children of I in column order.  The metadata of the next child (its fine row, weight, parent position and T-row
    // range: a chain of two dependent loads) is fetched while the T row of the current child is being accumulated.
    const int e0 = __ldg(r_off + I), e1 = __ldg(r_off + I + 1);
    int i_n = 0, q_n = 0, t_n = 0, t1_n = 0;
    float w_n = 0.f;
    if (e0 < e1) {
      i_n = __ldg(r_col + e0); w_n = __ldg(r_w + e0); q_n = ut_q[e0];
      t_n = __ldg(t_ptr + i_n); t1_n = __ldg(t_ptr + i_n + 1);
    }
    for (int e = e0; e < e1; ++e) {
      const float w = w_n;
      const int q = q_n, t1 = t1_n;
      int t = t_n;
      if (e + 1 < e1) {
        i_n = __ldg(r_col + e + 1); w_n = __ldg(r_w + e + 1); q_n = ut_q[e + 1];
        t_n = __ldg(t_ptr + i_n); t1_n = __ldg(t_ptr + i_n + 1);
      }
      // the slots of one T row are distinct: batches of 8 keep 16 independent global loads in flight per thread
      const unsigned char* cs = c_slot8 + (size_t)t * 8 + q;
      const float* tv_p = t_val + (size_t)t * 9 + k;
      for (; t + 8 <= t1; t += 8, cs += 64, tv_p += 72) {
        int sl[8];
        float tv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sl[u] = (int)__ldg(cs + u * 8);
          tv[u] = __ldg(tv_p + u * 9);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[sl[u] * GA_THREADS] += w * tv[u];
      }
      for (; t < t1; ++t)
        acc[(int)__ldg(c_slot8 + (size_t)t * 8 + q) * GA_THREADS] += w * __ldg(t_val + (size_t)t * 9 + k);
    }
  }
  __syncthreads();
  if (!live) return;
  const int c0 = __ldg(c_ptr + I), cn = __ldg(c_ptr + I + 1) - c0;
  for (int sl = k; sl < cn; sl += 9) {
    float m[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) m[e] = s_acc[sl * GA_THREADS + rl * 9 + e];
    if (spd) spd_block(m, spd_value);
    float* d = out + (size_t)__ldg(c_slot + c0 + sl) * 9;
#pragma unroll
    for (int e = 0; e < 9; ++e) d[e] = m[e];
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

int femb_set_pdl(int enable) {
  const int old = !g_no_pdl;
  g_no_pdl = !enable;
  return old;
}

int femb_set_gs_threads_per_row(int subs) {
  const int old = g_gs_subs;
  if (subs == 4 || subs == 8) g_gs_subs = subs;
  return old;
}

int femb_set_gs_rows_per_cta(int rows) {
  const int old = g_gs_rows;
  if (rows == 8 || rows == 16 || rows == 32) g_gs_rows = rows;
  return old;
}

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
  if (iterations == 0) return FEMB_OK;
  RowArgs a = base_args(lvl, values);
  a.parts = FEMB_PART_L | FEMB_PART_U; a.x = dx; a.y = dx; a.b = b; a.not_spd = not_spd;
  // (A cooperative single-launch variant -- grid-wide barrier between the 16 colour phases -- was measured slower than these
  // PDL colour launches at every level size, 375 vs 279 us at 132 651 rows and 189 vs 133 us at 17 576 rows for 3 iterations,
  // and was removed: DESIGN.md 4a.)
  bool first = true;                 // the first colour follows the memset (and whatever wrote the matrix): ordinary launch
  for (int it = 0; it < iterations; ++it) {
    for (int c = 7; c >= 0; --c) {                                                // backward sweep, object.py:893-898
      a.row_begin = lvl->color_start[c]; a.row_end = lvl->color_start[c + 1];
      // first half-sweep of the first iteration: dx == 0, rows of colour c only see higher colours
      a.parts = (it == 0) ? FEMB_PART_U : (FEMB_PART_L | FEMB_PART_U);
      int rc = launch_gs_color(a, !first, st);
      if (rc) return rc;
      first = false;
    }
    a.parts = FEMB_PART_L | FEMB_PART_U;
    for (int c = 0; c < 8; ++c) {                                                 // forward sweep, object.py:903-908
      a.row_begin = lvl->color_start[c]; a.row_end = lvl->color_start[c + 1];
      int rc = launch_gs_color(a, true, st);
      if (rc) return rc;
    }
  }
  return FEMB_OK;
}

int femb_gs_colors(const FembLevel* lvl, const float* values, const float* b, float* dx, int first_color, int last_color,
                   int parts, int* not_spd, void* stream) {
  FEMB_REQUIRE(lvl && values && b && dx, "null pointer");
  FEMB_REQUIRE(first_color >= 0 && first_color < 8 && last_color >= 0 && last_color < 8, "colour range");
  FEMB_REQUIRE(parts == FEMB_PART_U || parts == FEMB_PART_L || parts == (FEMB_PART_L | FEMB_PART_U), "parts");
  RowArgs a = base_args(lvl, values);
  a.parts = parts; a.x = dx; a.y = dx; a.b = b; a.not_spd = not_spd;
  const int step = first_color <= last_color ? 1 : -1;
  for (int c = first_color;; c += step) {
    a.row_begin = lvl->color_start[c]; a.row_end = lvl->color_start[c + 1];
    int rc = launch_gs_color(a, c != first_color, as_stream(stream));   // the first colour follows a halo exchange
    if (rc) return rc;
    if (c == last_color) break;
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

int femb_galerkin(const FembLevel* fine, const FembLevel* coarse, const float* fine_values, const FembTransfer* ut_hat,
                  const FembTransfer* us_hat, const FembGalerkinPlan* plan, float* t_values, int spd, float spd_value,
                  float* coarse_values, void* stream) {
  FEMB_REQUIRE(fine && coarse && fine_values && ut_hat && us_hat && plan && t_values && coarse_values, "null pointer");
  FEMB_REQUIRE(plan->n_fine == fine->n_vert && us_hat->nrow == fine->n_vert && ut_hat->nrow == plan->n_coarse &&
                   coarse->n_vert == plan->n_coarse, "inconsistent sizes");
  FEMB_REQUIRE(plan->t_slot && plan->c_slot8 && plan->ut_q && plan->t_ptr && plan->c_ptr && plan->c_slot, "plan not initialised");
  static bool configured_dev[64] = {};                         // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    cudaFuncSetAttribute(galerkin_ap_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GaSmem));
    cudaFuncSetAttribute(galerkin_ap_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GaSmem));
    configured = true;
  }
  cudaStream_t st = as_stream(stream);
  if (plan->n_fine > 0) {
    RowArgs a = base_args(fine, fine_values);
    if (plan->dyadic)
      galerkin_ap_kernel<true><<<cdiv(plan->n_fine, GA_ROWS), GA_THREADS, sizeof(GaSmem), st>>>(
          a, plan->n_fine, us_hat->offsets, us_hat->weights, plan->t_slot, plan->t_ptr, t_values);
    else
      galerkin_ap_kernel<false><<<cdiv(plan->n_fine, GA_ROWS), GA_THREADS, sizeof(GaSmem), st>>>(
          a, plan->n_fine, us_hat->offsets, us_hat->weights, plan->t_slot, plan->t_ptr, t_values);
    count_launch();
  }
  if (plan->n_coarse > 0) {
    galerkin_rt_kernel<<<cdiv(plan->n_coarse, GA_ROWS), GA_THREADS, 0, st>>>(
        plan->n_coarse, coarse->v2i, ut_hat->offsets, ut_hat->columns, ut_hat->weights, plan->ut_q, plan->t_ptr, t_values,
        plan->c_slot8, plan->c_ptr, plan->c_slot, spd, spd_value, coarse_values);
    count_launch();
  }
  return check_launch("galerkin");
}

int femb_adam_update(int n, const float* grad, float* m, float* v, float* m_hat, float* v_hat, float* x, float beta1, float beta2,
                     float bias1, float bias2, float lr, float epsilon, void* stream) {
  FEMB_REQUIRE(n >= 0 && grad && m && v && m_hat && v_hat && x, "null pointer or negative size");
  if (n == 0) return FEMB_OK;
  adam_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(n, grad, m, v, m_hat, v_hat, x, 1.f - beta1, beta1, 1.f - beta2, beta2,
                                                          bias1, bias2, lr, epsilon);
  count_launch();
  return check_launch("adam_kernel");
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
