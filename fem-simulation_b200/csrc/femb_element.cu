// Element physics of the 8-node trilinear hexahedron with a St. Venant-Kirchhoff material:
// rest data, force, 24x24 element Hessian, energy -- and the atomics-free assembly gathers.
//
// Reference (paths relative to /root/reference/exp1/quasi_simulation/):
//   prepare_kernal                      quai_hexagon.py:4-18
//   compute_elastic_energy              quai_hexagon.py:28-55
//   compute_elastic_hessian             quai_hexagon.py:69-106   (N_h*64 threads, 4608 float atomics / hex)
//   compute_partial_elastic_energy_X    quai_hexagon.py:109-133  (N_h*8 threads, 24 float atomics / thread)
//   gravity / pin-spring forces         quai_hexagon.py:160-186
//
// B200 design.  A thread carries TWO hexahedra, one in each half of 64-bit register pairs, and the whole element
// arithmetic runs on packed fp32 instructions (fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2: two FMAs per lane at the
// issue cost of one FFMA, scalar operands broadcast for free).  8 lanes work on a pair of hexahedra (4 pairs per warp).
// In phase A lane q evaluates everything that depends only on the Gauss point q
// (F, S, F F^T, g_b = J^-T dN_b, u_b = F g_b) once and parks it in shared memory; in phase B lane a accumulates, over the
// 8 Gauss points, the 5 blocks K[a][(a+k) & 7], k = 0..4 (the element matrix is symmetric: the cyclic half is enough)
// with the closed form
//     K_ab[j][i] = det * ( (g_a . S g_b) d_ji + mu (g_a . g_b)(F F^T)_ji + mu u_b[j] u_a[i] + la u_a[j] u_b[i] )
// (== d(P g_a det)_j / d x_b,i of quai_hexagon.py:86-106).  The half rows go to a scratch array and are summed into the
// fixed [L;D;U] value layout by a gather in a fixed order (no atomics):
//   * lattice levels (every level of a voxel mesh): the scratch is plane-major and blocked by 32 hexahedra,
//     EM[e >> 5][(a*5+k)*9+c][e & 31] (a block = the 46 080-byte tile one CTA of the element kernel produces and ships with
//     three TMA tensor stores), and fold_lattice_kernel gives every (row, block component) its own thread, which adds its
//     <= 64 contributions through compile-time offsets (the lattice offset of node b seen from node a) from one base
//     pointer per local node -- coalesced 128-byte lines, no plan decoding;
//   * any other hexahedral mesh: 192-byte half rows per (hexahedron, node) and the per-row fold plan, the precomputed
//     inverse of the element->nnz map (assemble_rows_kernel).
#include <cstdio>
#include <cstring>

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include "femb_common.cuh"

namespace femb {

constexpr int EPAIRS = FEMB_ASM_CHUNK_HEXES / 2;     // pairs of hexahedra per CTA
constexpr int EPB = FEMB_ASM_CHUNK_HEXES;            // hexahedra per CTA
constexpr int ETHREADS = EPAIRS * 8;                 // 128 threads

// ---- packed fp32: the two halves of a pair are two hexahedra -----------------------------------------------------
struct V2 { float2 v; };
__device__ __forceinline__ unsigned long long pk(const V2& a) { return *reinterpret_cast<const unsigned long long*>(&a.v); }
__device__ __forceinline__ V2 vfma(V2 a, V2 b, V2 c) {
  V2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d.v)) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return d;
}
__device__ __forceinline__ V2 vmul(V2 a, V2 b) {
  V2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d.v)) : "l"(pk(a)), "l"(pk(b)));
  return d;
}
__device__ __forceinline__ V2 vadd(V2 a, V2 b) {
  V2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d.v)) : "l"(pk(a)), "l"(pk(b)));
  return d;
}
__device__ __forceinline__ V2 vsub(V2 a, V2 b) {
  V2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d.v)) : "l"(pk(a)), "l"(pk(b)));
  return d;
}
__device__ __forceinline__ V2 vneg(V2 a) { V2 d; d.v = make_float2(-a.v.x, -a.v.y); return d; }
// the scalar instantiation rounds exactly like one half of the packed one (no contraction the packed code does not have)
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float vsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float vneg(float a) { return -a; }
template <class T> __device__ __forceinline__ T bc(float s);
template <> __device__ __forceinline__ float bc<float>(float s) { return s; }
template <> __device__ __forceinline__ V2 bc<V2>(float s) { V2 d; d.v = make_float2(s, s); return d; }
template <class T> __device__ __forceinline__ T dot3(T a0, T b0, T a1, T b1, T a2, T b2) { return vfma(a2, b2, vfma(a1, b1, vmul(a0, b0))); }
__device__ __forceinline__ V2 vshfl_xor(V2 a, int m) {
  V2 d;
  d.v.x = __shfl_xor_sync(0xffffffffu, a.v.x, m);
  d.v.y = __shfl_xor_sync(0xffffffffu, a.v.y, m);
  return d;
}
__device__ __forceinline__ void store2(float*, float, float) {}   // the scalar instantiation never writes records
__device__ __forceinline__ void store2(V2* p, V2 a, V2 b) { *reinterpret_cast<float4*>(p) = make_float4(a.v.x, a.v.y, b.v.x, b.v.y); }

// per (pair of hexahedra, Gauss point) record in shared memory, in pairs (T = V2), written and read as float4 = 2 pairs:
//   [0..5] det*mu * F F^T (symmetric: 00 01 02 11 12 22)   [6] det*mu   [7] det*la   [8..13] det * S (00 01 02 11 12 22)
//   [14 + 6 b ..] per node b: g_b[3], u_b[3]
constexpr int QP_FFT = 0;
constexpr int QP_C = 6;
constexpr int QP_S = 8;
constexpr int QP_GU = 14;
constexpr int QP_SIZE = 62;   // 31 float4: the 16-byte record writes of 8 consecutive lanes hit distinct banks (31 is odd)
// pairs between the records of consecutive hexahedron pairs: (8*62 + 4) / 2 = 250 float4 = 2 (mod 8), so the broadcast reads
// of the 4 pairs of a warp hit distinct banks
constexpr int PAIR_STRIDE = 8 * QP_SIZE + 4;
constexpr int XS_STRIDE = 52;                    // floats of staged node positions per pair (24 pairs + 4 pad: 20 mod 32)
constexpr int EK_ROW = FEMB_ELEM_K_FLOATS / 8;   // floats per stored row of an element matrix, general layout (5 blocks + 3 pad)
constexpr int EM_PLANES = 360;                   // lattice layout: planes (a*5+k)*9+c, blocked by 32 hexahedra: EM[e >> 5][plane][e & 31]
constexpr int EM_BLOCK = EM_PLANES * 32;         // floats per block of 32 hexahedra (46 080 bytes, written by one CTA of the element kernel)
constexpr int EM_STG = 36;                       // floats per plane in the staging area of a CTA (32 hexahedra + 4: see below)
constexpr int EM_TMA_BOX_ROWS = 120;             // TMA store variant: planes per tensor store (120 * 128 bytes keeps every box 1024-byte aligned)

struct ElemIn {
  const int* hex;
  const float* x;
  const float* invJ;
  const float* detJ;
  int e_begin, n_hex;    // the launch covers hexahedra [e_begin, n_hex); e_begin is a multiple of FEMB_ASM_CHUNK_HEXES
  int prefetch_ctas;     // > 0: every CTA pulls the vertex indices of the CTA that far ahead into L2
  float mu, la;
};

// shape-function gradients at Gauss point q (a runtime value: the lane)
__device__ __forceinline__ void shape_grads(int q, float (&dN)[8][3]) {
#pragma unroll
  for (int a = 0; a < 8; ++a) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = ((a >> (2 - c)) & 1) ? 1.0f : -1.0f;
      const int d0 = (c == 0) ? 1 : 0, d1 = (c == 2) ? 1 : 2;
      const float f0 = (((a >> (2 - d0)) & 1) == ((q >> (2 - d0)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
      const float f1 = (((a >> (2 - d1)) & 1) == ((q >> (2 - d1)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
      dN[a][c] = (s * f0) * f1 * 0.125f;
    }
  }
}

// Phase A at one Gauss point: the record (kRec), the nodal forces -(P g_b) det (kForce), the energy density times det
// (kEnergy).  T = V2: two hexahedra at once; T = float: one.
template <class T, bool kRec, bool kForce, bool kEnergy>
__device__ __forceinline__ void gauss_point(const T (&X)[8][3], const float (&dN)[8][3], const T (&iJ)[9], T det, float mu, float la,
                                            T* __restrict__ rec, T (&fq)[8][3], T& energy) {
  // F0 = sum_a x_a (x) dN_a ; F = F0 * invJ                                  (quai_hexagon.py:77-84)
  T F[9];
  {
    T F0[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        T v = vmul(X[0][r], bc<T>(dN[0][c]));
#pragma unroll
        for (int a = 1; a < 8; ++a) v = vfma(X[a][r], bc<T>(dN[a][c]), v);
        F0[3 * r + c] = v;
      }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) F[3 * r + c] = dot3(F0[3 * r], iJ[c], F0[3 * r + 1], iJ[3 + c], F0[3 * r + 2], iJ[6 + c]);
  }
  // E = 0.5 (F^T F - I), S = 2 mu E + la tr(E) I   (symmetric: 00 01 02 11 12 22)
  const T half = bc<T>(0.5f), one = bc<T>(1.f);
  T E[6];
  E[0] = vmul(half, vsub(dot3(F[0], F[0], F[3], F[3], F[6], F[6]), one));
  E[1] = vmul(half, dot3(F[0], F[1], F[3], F[4], F[6], F[7]));
  E[2] = vmul(half, dot3(F[0], F[2], F[3], F[5], F[6], F[8]));
  E[3] = vmul(half, vsub(dot3(F[1], F[1], F[4], F[4], F[7], F[7]), one));
  E[4] = vmul(half, dot3(F[1], F[2], F[4], F[5], F[7], F[8]));
  E[5] = vmul(half, vsub(dot3(F[2], F[2], F[5], F[5], F[8], F[8]), one));
  const T trE = vadd(vadd(E[0], E[3]), E[5]);
  if (kEnergy) {                                                             // quai_hexagon.py:47-53
    const T off = vfma(E[4], E[4], vfma(E[2], E[2], vmul(E[1], E[1])));
    const T sumE2 = vfma(bc<T>(2.f), off, dot3(E[0], E[0], E[3], E[3], E[5], E[5]));
    energy = vmul(vfma(sumE2, bc<T>(mu), vmul(bc<T>(0.5f * la), vmul(trE, trE))), det);
  }
  const T tm = bc<T>(2.f * mu), lt = vmul(bc<T>(la), trE);
  T S[6];
  S[0] = vfma(tm, E[0], lt); S[1] = vmul(tm, E[1]); S[2] = vmul(tm, E[2]);
  S[3] = vfma(tm, E[3], lt); S[4] = vmul(tm, E[4]); S[5] = vfma(tm, E[5], lt);
  T P[9];                                                                    // P = F S
  if (kForce) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      P[3 * r + 0] = dot3(F[3 * r], S[0], F[3 * r + 1], S[1], F[3 * r + 2], S[2]);
      P[3 * r + 1] = dot3(F[3 * r], S[1], F[3 * r + 1], S[3], F[3 * r + 2], S[4]);
      P[3 * r + 2] = dot3(F[3 * r], S[2], F[3 * r + 1], S[4], F[3 * r + 2], S[5]);
    }
  }
  if (kRec) {
    const T cmu = vmul(det, bc<T>(mu));
    store2(rec + QP_FFT + 0, vmul(cmu, dot3(F[0], F[0], F[1], F[1], F[2], F[2])), vmul(cmu, dot3(F[0], F[3], F[1], F[4], F[2], F[5])));
    store2(rec + QP_FFT + 2, vmul(cmu, dot3(F[0], F[6], F[1], F[7], F[2], F[8])), vmul(cmu, dot3(F[3], F[3], F[4], F[4], F[5], F[5])));
    store2(rec + QP_FFT + 4, vmul(cmu, dot3(F[3], F[6], F[4], F[7], F[5], F[8])), vmul(cmu, dot3(F[6], F[6], F[7], F[7], F[8], F[8])));
    store2(rec + QP_C, cmu, vmul(det, bc<T>(la)));
    store2(rec + QP_S + 0, vmul(det, S[0]), vmul(det, S[1]));
    store2(rec + QP_S + 2, vmul(det, S[2]), vmul(det, S[3]));
    store2(rec + QP_S + 4, vmul(det, S[4]), vmul(det, S[5]));
  }
  const T ndet = vneg(det);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    // g_b = invJ^T dN_b  (dF row of quai_hexagon.py:95)
    T g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = dot3(bc<T>(dN[b][0]), iJ[j], bc<T>(dN[b][1]), iJ[3 + j], bc<T>(dN[b][2]), iJ[6 + j]);
    // nodal force contribution  -(P g_b) det                                 (quai_hexagon.py:128-133)
    if (kForce) {
#pragma unroll
      for (int r = 0; r < 3; ++r) fq[b][r] = vmul(ndet, dot3(P[3 * r], g[0], P[3 * r + 1], g[1], P[3 * r + 2], g[2]));
    }
    if (kRec) {
      T u[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) u[r] = dot3(F[3 * r], g[0], F[3 * r + 1], g[1], F[3 * r + 2], g[2]);
      T* o = rec + QP_GU + 6 * b;
      store2(o, g[0], g[1]); store2(o + 2, g[2], u[0]); store2(o + 4, u[1], u[2]);
    }
  }
}
// ---- row-centric assembly of one row (vertex v, index-order row r) by one warp -------------------------
// The row needs row `a` of the element matrices of its <= 8 incident hexahedra: 8 blocks = 288 contiguous,
// 16-byte aligned bytes each, fetched as 18 float4 by 18 lanes (all 8 fetches in flight before the first is
// consumed).  The blocks are folded into the row's <= 27 value slots through the element->nnz map in shared
// memory (fixed order: incidence order, then b; no atomics) and the row's L, D and U segments are written
// with coalesced stores.
constexpr int AS_SLOTS = FEMB_FOLD_SLOTS;       // 27: blocks of a row of a hexahedral lattice
struct AsmArgs {
  int n_vert;
  const int* i2v; const int* v2i; const int* vert_elem; const int* elem2nnz; const int* L_ptr; const int* U_ptr;
  const unsigned char* fold;     // [n_vert][27][8] fold plan (femb_build_fold_plan)
  int off_l, off_d, off_u;
  const float* elem_k; const float* diag_add; float* values;
};
// Stored layout of a half row (EK_ROW = 48 floats): positions 0..4 as 8 floats each (elements 0..7 of the block, 16-byte
// aligned) at [8p, 8p+8), their ninth elements ("tails") at [40, 45), 3 pad floats.  The staged copy of a row in shared
// memory keeps that shape per incident hexahedron j: 8 positions x 8 floats, then 8 tails; positions 5..7 are the
// mirrored blocks as they are stored (position m of row (a-m)&7) -- they are transposed in registers by the fold.
constexpr int AS_JSTRIDE = 84;                    // 64 + 8 own tails + 12 mirror tails; = 20 (mod 32) spreads the hexahedra over the banks
constexpr int AS_TAIL = 64;                       // tails of positions 0..4 (two float4 as stored)
constexpr int AS_MTAIL = 72;                      // tail float4 of rows (a-m)&7, m = 1..3, as stored: tail of position 8-m at +4(m-1)+m
constexpr int AS_ZERO = 8 * AS_JSTRIDE;           // a block of zeros (+ zero tail at +8): what an absent contributor reads
struct AsmWarpSmem {
  alignas(16) float stage[8 * AS_JSTRIDE + 12];   // the folded row (27 x 9 floats) reuses the front of it
};
// per-row record of the fold plan: 28 x 4 contributor bytes, then a 16-byte header
constexpr int FP_ROW = FEMB_FOLD_PLAN_BYTES;       // 128
// what a warp needs to know about a row before touching the element matrices; one coalesced read of the plan record
// (+ the incidence list), prefetched one row ahead
struct AsmMeta {
  unsigned plan;     // lane s < 27: the <= 4 contributors of block s (the diagonal block: its first 4); lane 27: contributors
                     // 5..8 of the diagonal block; one byte each, j*8+p, 255 = none
  int my_ea;         // lane j < 8: j-th incident hexahedron*8 + local node, -1 = none
  int r, lp, up, nl, nu;
};
__device__ __forceinline__ AsmMeta asm_meta(const AsmArgs& A, int v, int lane) {
  AsmMeta M;
  // the whole record is one 128-byte line: 28 plan words, then {row, L_ptr[row], U_ptr[row], nL | nU << 8}
  const unsigned w = __ldg(reinterpret_cast<const unsigned*>(A.fold + (size_t)v * FP_ROW) + lane);
  M.plan = lane < AS_SLOTS + 1 ? w : 0xffffffffu;
  M.my_ea = lane < 8 ? __ldg(A.vert_elem + (size_t)v * 8 + lane) : -1;
  M.r = (int)__shfl_sync(0xffffffffu, w, 28);
  M.lp = (int)__shfl_sync(0xffffffffu, w, 29);
  M.up = (int)__shfl_sync(0xffffffffu, w, 30);
  const unsigned nn = __shfl_sync(0xffffffffu, w, 31);
  M.nl = (int)(nn & 255u); M.nu = (int)((nn >> 8) & 255u);
  return M;
}
// what depends only on the lane: computed once per warp, used for every row the warp folds.  Per incident hexahedron one
// LDG.128 per lane fetches: lanes 0..11 the 12 float4 of the own half row; lanes 12..17 the two float4 of position m of
// row (a-m)&7 (m = 1..3); lanes 18..20 the tail float4 of those rows.
struct AsmLane {
  int nb_m;        // 0: own row, m: row (a - m) & 7
  int src4;        // float4 index inside that row
  int dst;         // float offset inside the staged hexahedron where the float4 goes
};
__device__ __forceinline__ AsmLane asm_lane(int lane) {
  AsmLane L;
  L.nb_m = 0; L.src4 = 0; L.dst = 0;
  if (lane < 10) { L.src4 = lane; L.dst = 4 * lane; }                                   // positions 0..4, elements 0..7
  else if (lane < 12) { L.src4 = lane; L.dst = AS_TAIL + 4 * (lane - 10); }              // tails of positions 0..4 (+ pad)
  else if (lane < 18) { const int m = (lane - 12) / 2 + 1, h = (lane - 12) & 1; L.nb_m = m; L.src4 = 2 * m + h; L.dst = (8 - m) * 8 + 4 * h; }
  else if (lane < 21) { const int m = lane - 17; L.nb_m = m; L.src4 = 10; L.dst = AS_MTAIL + 4 * (m - 1); }
  return L;
}
// float offset of the ninth element of rotated position p inside a staged hexahedron
__device__ __forceinline__ int staged_tail(int p) { return p < 5 ? AS_TAIL + p : AS_MTAIL + 36 - 5 * p; }   // m = 8-p: 4(m-1)+m
// element k of the block at rotated position p of staged hexahedron j (mirrored positions are stored transposed)
__device__ __forceinline__ float staged_elem(const float* stage, int j, int p, int k) {
  const int ks = p >= 5 ? 3 * (k % 3) + k / 3 : k;
  return ks < 8 ? stage[j * AS_JSTRIDE + p * 8 + ks] : stage[j * AS_JSTRIDE + staged_tail(p)];
}

// rows with more than 27 blocks (not produced by hexahedral lattices): one slot at a time through elem2nnz
__device__ __noinline__ void assemble_row_slow(const int* elem2nnz, const float* diag_add, float* values, int off_l, int off_d,
                                               int off_u, int r, int lane, int my_ea, int lp, int nl, int up, int nu,
                                               const float* stage) {
  for (int part = 0; part < 3; ++part) {
    const int cnt = part == 0 ? nl : (part == 1 ? 1 : nu);
    for (int li = 0; li < cnt; ++li) {
      const int slot = part == 0 ? off_l + lp + li : (part == 1 ? off_d + r : off_u + up + li);
      float sum = (part == 1 && diag_add != nullptr && (lane == 0 || lane == 4 || lane == 8)) ? __ldg(diag_add + r) : 0.f;
      for (int t = 0; t < 64; ++t) {
        const int ea = __shfl_sync(0xffffffffu, my_ea, t >> 3);
        if (ea < 0) continue;
        if (__ldg(elem2nnz + (size_t)ea * 8 + (t & 7)) == slot && lane < 9)
          sum += staged_elem(stage, t >> 3, ((t & 7) - (ea & 7)) & 7, lane);              // rotated position of column node b
      }
      if (lane < 9) values[(size_t)slot * 9 + lane] = sum;
    }
  }
}

__device__ __forceinline__ void assemble_row(const AsmArgs& A, const AsmLane& LN, const AsmMeta& M, int lane, AsmWarpSmem& sm) {
  const int r = M.r, lp = M.lp, nl = M.nl, up = M.up, nu = M.nu, my_ea = M.my_ea;
  const unsigned plan = M.plan;
  // One LDG.128 + one STS.128 per lane and incident hexahedron.  A missing hexahedron (ea < 0, boundary vertex) is read
  // from hexahedron 0 instead: no contributor of the plan points at it, so what lands in its stage is never looked at.
  if (lane < 21) {
    float4 seg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ea = max(__shfl_sync(0x001fffffu, my_ea, j), 0);
      const int rowid = (ea & ~7) | ((ea - LN.nb_m) & 7);
      seg[j] = __ldg(reinterpret_cast<const float4*>(A.elem_k + (size_t)rowid * EK_ROW) + LN.src4);
    }
    float* d = sm.stage + LN.dst;
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(d + j * AS_JSTRIDE) = seg[j];
  }
  __syncwarp();
  const unsigned first = __shfl_sync(0xffffffffu, plan, 0) & 255u;
  if (first == 254u) {                                       // oversized row (plan not applicable); warp-uniform
    assemble_row_slow(A.elem2nnz, A.diag_add, A.values, A.off_l, A.off_d, A.off_u, r, lane, my_ea, lp, nl, up, nu, sm.stage);
    __syncwarp();
    return;
  }
  // lane s folds block s: contributors in incidence order; a mirrored contributor (position >= 5) is added transposed.
  // Two vertices share at most 4 hexahedra, so only the diagonal block has more than 4 contributors: lane 27 folds its
  // second half and hands it over by shuffle.
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned idx = (plan >> (8 * c)) & 255u;
    const bool have = idx < 64u;                             // no branch: an absent contributor reads the zero block
    const unsigned j = idx >> 3, p = idx & 7u;
    const unsigned off = have ? j * AS_JSTRIDE + p * 8 : (unsigned)AS_ZERO;
    const unsigned toff = have ? j * AS_JSTRIDE + (p < 5u ? AS_TAIL + p : AS_MTAIL + 36u - 5u * p) : (unsigned)AS_ZERO + 8;
    const bool tr = have && p >= 5u;
    const float4* s4 = reinterpret_cast<const float4*>(sm.stage + off);
    const float4 b0 = s4[0], b1 = s4[1];
    const float t8 = sm.stage[toff];
    acc[0] += b0.x; acc[4] += b1.x; acc[8] += t8;
    acc[1] += tr ? b0.w : b0.y; acc[3] += tr ? b0.y : b0.w;
    acc[2] += tr ? b1.z : b0.z; acc[6] += tr ? b0.z : b1.z;
    acc[5] += tr ? b1.w : b1.y; acc[7] += tr ? b1.y : b1.w;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float hi = __shfl_sync(0xffffffffu, acc[k], AS_SLOTS);   // second half of the diagonal block
    if (lane == nl) acc[k] += hi;
  }
  __syncwarp();                                              // every lane is done reading the staged blocks
  float* out = sm.stage;                                     // the folded row: L blocks, D, U blocks
  if (lane < AS_SLOTS) {
#pragma unroll
    for (int k = 0; k < 9; ++k) out[lane * 9 + k] = acc[k];
  }
  __syncwarp();
  float* outL = A.values + ((size_t)A.off_l + lp) * 9;
  for (int i = lane; i < nl * 9; i += 32) outL[i] = out[i];
  if (lane < 9) {
    float d = out[nl * 9 + lane];
    if (A.diag_add != nullptr && (lane == 0 || lane == 4 || lane == 8)) d += __ldg(A.diag_add + r);
    A.values[((size_t)A.off_d + r) * 9 + lane] = d;
  }
  float* outU = A.values + ((size_t)A.off_u + up) * 9;
  for (int i = lane; i < nu * 9; i += 32) outU[i] = out[(nl + 1) * 9 + i];
  __syncwarp();                                              // the warp's scratch is reused by its next row
}

// setup: the fold plan of every row (vertex order).  One warp per vertex.
__global__ void __launch_bounds__(256) fold_plan_kernel(AsmArgs A, unsigned char* __restrict__ fold) {
  __shared__ int s_loc[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int v = blockIdx.x * 8 + w;
  if (v >= A.n_vert) return;
  const int r = __ldg(A.v2i + v);
  const int lp = __ldg(A.L_ptr + r), nl = __ldg(A.L_ptr + r + 1) - lp;
  const int up = __ldg(A.U_ptr + r), nu = __ldg(A.U_ptr + r + 1) - up;
  const int my_ea = lane < 8 ? __ldg(A.vert_elem + (size_t)v * 8 + lane) : -1;
  unsigned char* out = fold + (size_t)v * FP_ROW;
  if (lane == 0) {                                            // header: r, lp, up, nl + (nu << 8)
    int* h = reinterpret_cast<int*>(out + 28 * 4);
    h[0] = r; h[1] = lp; h[2] = up; h[3] = (nl & 255) | ((nu & 255) << 8);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = lane + 32 * h, j = t >> 3, b = t & 7;
    const int ea = __shfl_sync(0xffffffffu, my_ea, j);
    int loc = -1;
    if (ea >= 0) {
      const int s = __ldg(A.elem2nnz + (size_t)ea * 8 + b);
      loc = s < A.off_d ? (s - A.off_l - lp) : (s < A.off_u ? nl : nl + 1 + (s - A.off_u - up));
    }
    // the element kernel stores row a of an element matrix rotated: column node b sits at position (b - a) & 7
    s_loc[w][(t & ~7) | ((b - (ea & 7)) & 7)] = loc;
  }
  __syncwarp();
  // lane s < 27: contributors of block s; more than 4 are allowed only for the diagonal block (slot nl), whose
  // contributors 5..8 go to lane 27.  Anything else (or a row with more than 27 blocks) takes the slow path: 254.
  bool bad = nl + 1 + nu > AS_SLOTS;
  unsigned char mine[8];
  int n = 0;
  if (lane < AS_SLOTS && !bad) {
    for (int t = 0; t < 64; ++t)
      if (s_loc[w][t] == lane) {
        if (n < 8) mine[n] = (unsigned char)t;
        ++n;
      }
    if (n > (lane == nl ? 8 : 4)) bad = true;
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane < AS_SLOTS) {
    for (int c = 0; c < 4; ++c) out[lane * 4 + c] = bad ? 254 : (c < n ? mine[c] : 255);
  }
  // the diagonal lane also writes the overflow word of lane 27
  const int dn = __shfl_sync(0xffffffffu, n, bad ? 0 : nl);
  unsigned hi = 0xffffffffu;
  if (!bad && lane == nl) {
    hi = 0;
    for (int c = 0; c < 4; ++c) hi |= (unsigned)(4 + c < n ? mine[4 + c] : 255) << (8 * c);
  }
  hi = __shfl_sync(0xffffffffu, hi, bad ? 0 : nl);
  (void)dn;
  if (lane == 0) *reinterpret_cast<unsigned*>(out + 27 * 4) = bad ? 0xfefefefeu : hi;
}

// lattice offset slot of node b seen from node a: 9 (dx+1) + 3 (dy+1) + (dz+1)   (local node = 4 bx + 2 by + bz)
__host__ __device__ __forceinline__ constexpr int offset_slot(int a, int b) {
  return 13 + 9 * (((b >> 2) & 1) - ((a >> 2) & 1)) + 3 * (((b >> 1) & 1) - ((a >> 1) & 1)) + ((b & 1) - (a & 1));
}

// kMode: 0 = forces only, 1 = Hessian only, 2 = both.  kSoA: element matrices in the plane-major lattice layout.
// kLa: the Lame parameter lambda is non-zero (a compile-time switch: a run-time one inside the Gauss-point loop costs a
// register shuffle per accumulator around the branch)
// kTma (lattice layout only): the CTA's 360 x 32 tile leaves through ONE TMA tensor store instead of a read-back loop.
constexpr int RD_PITCH = 608;                    // bytes between the staged rest data of consecutive pairs (576 + 32: see element_kernel)
template <int kMode, bool kSoA, bool kLa, bool kTma, bool kBulk = false>
__global__ void __launch_bounds__(ETHREADS, kMode == 0 ? 5 : 3) element_kernel(ElemIn in, float* __restrict__ elem_f, float* __restrict__ elem_k,
                                                                                const __grid_constant__ CUtensorMap em_map) {
  constexpr bool kForce = (kMode != 1);
  constexpr bool kHess = (kMode != 0);
  extern __shared__ __align__(1024) float smem[];
  float* s_x = smem;                                                  // [EPAIRS][XS_STRIDE]: node positions of the pairs
  V2* s_rec = reinterpret_cast<V2*>(smem + EPAIRS * XS_STRIDE);       // [EPAIRS][PAIR_STRIDE]: 8 records per pair (Hessian only)
  const int sub = threadIdx.x & 7;
  const int lp = threadIdx.x >> 3;
  const int e0 = in.e_begin + (blockIdx.x * EPAIRS + lp) * 2, e1 = e0 + 1;   // this launch covers hexahedra [e_begin, n_hex)
  const bool live0 = e0 < in.n_hex, live1 = e1 < in.n_hex;
  const int ee0 = live0 ? e0 : in.n_hex - 1, ee1 = live1 ? e1 : in.n_hex - 1;   // dead halves shadow a valid hexahedron
  // Rest data through bulk copies (every CTA but a ragged last one), issued before anything else so that their latency
  // overlaps the chain hexahedron -> vertex positions below: the warp's 8 hexahedra are 2304 + 256 contiguous bytes of
  // invJ / detJ; they are copied into the head of the warp's own record area (dead until phase A writes it, after the
  // __syncwarp that follows the reads), one 576-byte copy per pair at a pitch of 608 bytes so that the per-lane reads (lane
  // stride 9 words, pair stride 152 words) are conflict-free, instead of 20 strided global loads per lane.  (Hessian kernels
  // only: the force-only kernel has no record area, and with an area of its own it spilled: 226 vs 199 us.)
  const bool bulk_cta = kBulk && in.e_begin + (int)(blockIdx.x + 1) * EPB <= in.n_hex;
  uint64_t* bar = nullptr;
  unsigned char* stg_w = nullptr;
  if (bulk_cta) {
    const int w = threadIdx.x >> 5;
    bar = reinterpret_cast<uint64_t*>(s_x + (w * 4) * XS_STRIDE + 48);                 // the 16 pad bytes of the warp's first position row
    stg_w = reinterpret_cast<unsigned char*>(s_rec + (size_t)(w * 4) * PAIR_STRIDE);
    if ((threadIdx.x & 31) == 0) {
      const int e_w = in.e_begin + (blockIdx.x * EPAIRS + w * 4) * 2;                   // first hexahedron of the warp
      mbar_init(bar, 1);
      mbar_expect_tx(bar, 4 * 576 + 256);
      bulk_g2s(stg_w + 4 * RD_PITCH, in.detJ + (size_t)e_w * 8, 256, bar);
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) bulk_g2s(stg_w + pr * RD_PITCH, in.invJ + (size_t)(e_w + 2 * pr) * 72, 576, bar);
    }
    __syncwarp();
  }
  if (in.prefetch_ctas > 0 && threadIdx.x < 8) {
    // the CTA's first dependent load chain (hexahedron -> vertex positions) starts with a DRAM miss; a CTA a quarter wave
    // ahead finds its 1 KB of vertex indices in L2 instead
    const long long ep = (long long)in.e_begin + ((long long)blockIdx.x + in.prefetch_ctas) * EPB;
    if (ep + EPB <= in.n_hex) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(in.hex + ep * 8) + threadIdx.x * 128));
  }
  // node positions: lane `sub` loads node `sub` of both hexahedra, the 8 lanes of the pair share them through shared memory
  {
    const int v0 = __ldg(in.hex + (size_t)ee0 * 8 + sub), v1 = __ldg(in.hex + (size_t)ee1 * 8 + sub);
    float2* d = reinterpret_cast<float2*>(s_x + lp * XS_STRIDE) + 3 * sub;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = make_float2(__ldg(in.x + 3 * (size_t)v0 + c), __ldg(in.x + 3 * (size_t)v1 + c));
  }
  float dN[8][3];
  shape_grads(sub, dN);
  V2 iJ[9], det;
  if (bulk_cta) {
    mbar_wait(bar, 0);
    const int lane = threadIdx.x & 31;
    const float* sj = reinterpret_cast<const float*>(stg_w + (lane >> 3) * RD_PITCH) + sub * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) iJ[k].v = make_float2(sj[k], sj[72 + k]);
    const float* sd = reinterpret_cast<const float*>(stg_w + 4 * RD_PITCH) + (lane >> 3) * 16 + sub;
    det.v = make_float2(sd[0], sd[8]);
  } else {
    const float* p0 = in.invJ + ((size_t)ee0 * 8 + sub) * 9;
    const float* p1 = in.invJ + ((size_t)ee1 * 8 + sub) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) iJ[k].v = make_float2(__ldg(p0 + k), __ldg(p1 + k));
    det.v = make_float2(__ldg(in.detJ + (size_t)ee0 * 8 + sub), __ldg(in.detJ + (size_t)ee1 * 8 + sub));
  }
  __syncwarp();
  {
    V2 X[8][3];
    const float4* xs4 = reinterpret_cast<const float4*>(s_x + lp * XS_STRIDE);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float4 t = xs4[i];
      X[(2 * i) / 3][(2 * i) % 3].v = make_float2(t.x, t.y);
      X[(2 * i + 1) / 3][(2 * i + 1) % 3].v = make_float2(t.z, t.w);
    }
    V2 fq[8][3], en;
    gauss_point<V2, kHess, kForce, false>(X, dN, iJ, det, in.mu, in.la, s_rec + (size_t)lp * PAIR_STRIDE + sub * QP_SIZE, fq, en);
    if (kForce) {
      // lane q holds the Gauss-point contributions to all 8 nodes; a reduce-scatter over the 8 lanes of the
      // pair (xor 4, 2, 1) leaves lane a with the force on node a
      V2 h4[4][3], h2[2][3], f[3];
      const bool up4 = (sub & 4) != 0, up2 = (sub & 2) != 0, up1 = (sub & 1) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const V2 give = up4 ? fq[i][c] : fq[i + 4][c], keep = up4 ? fq[i + 4][c] : fq[i][c];
          h4[i][c] = vadd(keep, vshfl_xor(give, 4));
        }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const V2 give = up2 ? h4[i][c] : h4[i + 2][c], keep = up2 ? h4[i + 2][c] : h4[i][c];
          h2[i][c] = vadd(keep, vshfl_xor(give, 2));
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const V2 give = up1 ? h2[0][c] : h2[1][c], keep = up1 ? h2[1][c] : h2[0][c];
        f[c] = vadd(keep, vshfl_xor(give, 1));
      }
      if (live0) {
        float* dst = elem_f + ((size_t)e0 * 8 + sub) * 3;
        dst[0] = f[0].v.x; dst[1] = f[1].v.x; dst[2] = f[2].v.x;
      }
      if (live1) {
        float* dst = elem_f + ((size_t)e1 * 8 + sub) * 3;
        dst[0] = f[0].v.y; dst[1] = f[1].v.y; dst[2] = f[2].v.y;
      }
    }
  }
  if (!kHess) return;
  __syncwarp();                                              // the 8 records of a pair are written by 8 lanes of one warp

  // ---- phase B: lane a accumulates the 5 blocks K[a][(a+k) & 7], k = 0..4 ------------------------------
  // K_ab[j][i] = sum_q  d_ji (t_a . g_b) + (det mu FF^T)_ji (g_a . g_b) + det mu u_b[j] u_a[i] + det la u_a[j] u_b[i]
  // with t_a = det S g_a (S symmetric: g_a . S g_b == S g_a . g_b); every term is one FFMA2 into an accumulator pair.
  // The element matrix is symmetric (K_ba = K_ab^T), so the cyclic half {(a, a+k)} is enough: the blocks
  // K[a][(a-m) & 7], m = 1..3, are the transposes of block m of row (a-m) & 7 and are picked up from there by the fold.
  V2 K[5][9], Ks[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    Ks[k] = bc<V2>(0.f);
#pragma unroll
    for (int i = 0; i < 9; ++i) K[k][i] = bc<V2>(0.f);
  }
  constexpr bool has_la = kLa;
  const V2* rp = s_rec + (size_t)lp * PAIR_STRIDE;
  // fully unrolled: a rolled loop pays a register copy per accumulator pair per trip (the packed ops are not in place)
#pragma unroll
  for (int q = 0; q < 8; ++q, rp += QP_SIZE) {
    const float4* r4 = reinterpret_cast<const float4*>(rp);
    V2 FFT[6], cmu, cla, Sd[6];
    {
      const float4 h0 = r4[0], h1 = r4[1], h2 = r4[2], h3 = r4[3], h4 = r4[4], h5 = r4[5], h6 = r4[6];   // broadcast reads
      FFT[0].v = make_float2(h0.x, h0.y); FFT[1].v = make_float2(h0.z, h0.w); FFT[2].v = make_float2(h1.x, h1.y);
      FFT[3].v = make_float2(h1.z, h1.w); FFT[4].v = make_float2(h2.x, h2.y); FFT[5].v = make_float2(h2.z, h2.w);
      cmu.v = make_float2(h3.x, h3.y); cla.v = make_float2(h3.z, h3.w);
      Sd[0].v = make_float2(h4.x, h4.y); Sd[1].v = make_float2(h4.z, h4.w); Sd[2].v = make_float2(h5.x, h5.y);
      Sd[3].v = make_float2(h5.z, h5.w); Sd[4].v = make_float2(h6.x, h6.y); Sd[5].v = make_float2(h6.z, h6.w);
    }
    V2 g_a[3], u_a[3];
    {
      const float4* na = reinterpret_cast<const float4*>(rp + QP_GU + 6 * sub);   // own node
      const float4 o0 = na[0], o1 = na[1], o2 = na[2];
      g_a[0].v = make_float2(o0.x, o0.y); g_a[1].v = make_float2(o0.z, o0.w); g_a[2].v = make_float2(o1.x, o1.y);
      u_a[0].v = make_float2(o1.z, o1.w); u_a[1].v = make_float2(o2.x, o2.y); u_a[2].v = make_float2(o2.z, o2.w);
    }
    V2 t_a[3];
    t_a[0] = dot3(Sd[0], g_a[0], Sd[1], g_a[1], Sd[2], g_a[2]);
    t_a[1] = dot3(Sd[1], g_a[0], Sd[3], g_a[1], Sd[4], g_a[2]);
    t_a[2] = dot3(Sd[2], g_a[0], Sd[4], g_a[1], Sd[5], g_a[2]);
    const V2 ua_mu[3] = {vmul(cmu, u_a[0]), vmul(cmu, u_a[1]), vmul(cmu, u_a[2])};
    V2 ua_la[3];
    if (has_la) {
#pragma unroll
      for (int i = 0; i < 3; ++i) ua_la[i] = vmul(cla, u_a[i]);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      V2 gb[3], ub[3];
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { gb[i] = g_a[i]; ub[i] = u_a[i]; }
      } else {
        const float4* nb = reinterpret_cast<const float4*>(rp + QP_GU + 6 * ((sub + k) & 7));
        const float4 p0 = nb[0], p1 = nb[1], p2 = nb[2];
        gb[0].v = make_float2(p0.x, p0.y); gb[1].v = make_float2(p0.z, p0.w); gb[2].v = make_float2(p1.x, p1.y);
        ub[0].v = make_float2(p1.z, p1.w); ub[1].v = make_float2(p2.x, p2.y); ub[2].v = make_float2(p2.z, p2.w);
      }
      Ks[k] = vfma(t_a[0], gb[0], vfma(t_a[1], gb[1], vfma(t_a[2], gb[2], Ks[k])));
      const V2 s2 = dot3(g_a[2], gb[2], g_a[1], gb[1], g_a[0], gb[0]);
      K[k][0] = vfma(FFT[0], s2, vfma(ub[0], ua_mu[0], K[k][0]));
      K[k][1] = vfma(FFT[1], s2, vfma(ub[0], ua_mu[1], K[k][1]));
      K[k][2] = vfma(FFT[2], s2, vfma(ub[0], ua_mu[2], K[k][2]));
      K[k][3] = vfma(FFT[1], s2, vfma(ub[1], ua_mu[0], K[k][3]));
      K[k][4] = vfma(FFT[3], s2, vfma(ub[1], ua_mu[1], K[k][4]));
      K[k][5] = vfma(FFT[4], s2, vfma(ub[1], ua_mu[2], K[k][5]));
      K[k][6] = vfma(FFT[2], s2, vfma(ub[2], ua_mu[0], K[k][6]));
      K[k][7] = vfma(FFT[4], s2, vfma(ub[2], ua_mu[1], K[k][7]));
      K[k][8] = vfma(FFT[5], s2, vfma(ub[2], ua_mu[2], K[k][8]));
      if (has_la) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) K[k][3 * j + i] = vfma(ua_la[j], ub[i], K[k][3 * j + i]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    K[k][0] = vadd(K[k][0], Ks[k]); K[k][4] = vadd(K[k][4], Ks[k]); K[k][8] = vadd(K[k][8], Ks[k]);
  }
  if (kSoA && kTma) {
    // Lattice layout through the TMA: the tile is written as 360 rows (plane R = a*45 + k*9 + i) of 32 hexahedra (128
    // bytes) in the 128-byte swizzle (16-byte chunk ^= R & 7; for a fixed (k, i) the 8 nodes of a pair hit 8 different
    // chunks because 45 is odd, so the 8-byte writes of a warp -- 4 pairs x 8 nodes -- take 2 wavefronts per 256 bytes),
    // which is the shared-memory image of boxes {32 hexahedra, 120 planes} of the 2-D view {hexahedron, plane} of the
    // scratch: three tensor stores per CTA, no read-back, no store loop.
    __syncthreads();
    const unsigned stg = (smem_u32(smem) + (unsigned)(EPAIRS * XS_STRIDE * sizeof(float)) + 1023u) & ~1023u;
    const unsigned rbase = stg + (unsigned)sub * (45u * 128u) + (unsigned)(lp & 1) * 8u;
    const unsigned r7 = (unsigned)(sub * 45) & 7u, ch = (unsigned)(lp >> 1);
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const unsigned addr = rbase + (unsigned)((k * 9 + i) * 128) + ((ch ^ ((r7 + (unsigned)(k * 9 + i)) & 7u)) << 4);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(K[k][i].v.x), "f"(K[k][i].v.y) : "memory");
      }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int blk = in.e_begin / EPB + (int)blockIdx.x;
#pragma unroll
      for (int j = 0; j < 3; ++j) tma_store_3d_saddr(&em_map, stg + (unsigned)(j * EM_TMA_BOX_ROWS * 128), 0, j * EM_TMA_BOX_ROWS, blk);
      tma_store_commit_and_wait_read();
    }
  } else if (kSoA) {
    // Lattice layout: plane (a*5+k)*9+c holds one float per hexahedron.  The CTA's 32 hexahedra are transposed through
    // the record area (dead now, but shared by all warps: two block barriers) and leave as 360 full 128-byte lines.
    // EM_STG = 36: the 8-byte writes of a half warp (2 pairs x 8 nodes, 45*36 = 20 (mod 32) floats apart) and the 16-byte
    // reads of a quarter warp (one plane) are conflict-free.
    __syncthreads();
    float* stg = reinterpret_cast<float*>(s_rec);
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int i = 0; i < 9; ++i) *reinterpret_cast<float2*>(stg + ((sub * 5 + k) * 9 + i) * EM_STG + 2 * lp) = K[k][i].v;
    __syncthreads();
    float4* gdst = reinterpret_cast<float4*>(elem_k + ((size_t)(in.e_begin / EPB) + blockIdx.x) * EM_BLOCK);   // the CTA's block: contiguous
    for (int idx = threadIdx.x; idx < EM_PLANES * 8; idx += ETHREADS) {
      const int j = idx >> 3, part = idx & 7;
      gdst[idx] = *reinterpret_cast<const float4*>(stg + j * EM_STG + 4 * part);
    }
  } else {
    // General layout: row a of the element matrix is stored ROTATED and HALF: position p = 0..4 holds the block of column
    // (a + p) & 7 (45 floats + 3 pad = 192 bytes: elements 0..7 of position p at [8p, 8p+8), the ninth elements at [40, 45)).
    // The blocks of columns (a - m) & 7, m = 1..3, are the transposes of position m of row (a - m) & 7 and are picked up from
    // there by the assembly.  The warp's 64 half rows are contiguous in elem_k (8 consecutive hexahedra x 8 rows x 192
    // bytes): they are transposed through the warp's own record area (dead now; read only by this warp) and stored as
    // fully coalesced 512-byte writes.
    constexpr int WROW = EK_ROW + 4;                                   // padded row: 16-byte stores of 8 lanes hit distinct banks
    const int lane = threadIdx.x & 31;
    float* wst = reinterpret_cast<float*>(s_rec + (size_t)(threadIdx.x >> 5) * 4 * PAIR_STRIDE);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float row[48];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) row[8 * k + i] = h ? K[k][i].v.y : K[k][i].v.x;
        row[40 + k] = h ? K[k][8].v.y : K[k][8].v.x;
      }
      row[45] = 0.f; row[46] = 0.f; row[47] = 0.f;
      float4* my = reinterpret_cast<float4*>(wst + (((lane >> 3) * 2 + h) * 8 + sub) * WROW);
#pragma unroll
      for (int v = 0; v < EK_ROW / 4; ++v) my[v] = make_float4(row[4 * v], row[4 * v + 1], row[4 * v + 2], row[4 * v + 3]);
    }
    __syncwarp();
    const int e_w = in.e_begin + (blockIdx.x * EPAIRS + (threadIdx.x >> 5) * 4) * 2;   // first hexahedron of the warp
    float4* gdst = reinterpret_cast<float4*>(elem_k + (size_t)e_w * 8 * EK_ROW);
#pragma unroll 4
    for (int i = 0; i < 2 * EK_ROW / 4; ++i) {
      const int g = i * 32 + lane;                                     // float4 index inside the warp's 12288 bytes
      const int rw = g / (EK_ROW / 4), c = g - rw * (EK_ROW / 4);
      if (e_w + (rw >> 3) < in.n_hex) gdst[g] = *reinterpret_cast<const float4*>(wst + rw * WROW + 4 * c);
    }
  }
}

static_assert(EPAIRS * XS_STRIDE * 4 + 1023 + EM_PLANES * 128 <= (EPAIRS * XS_STRIDE + EPAIRS * PAIR_STRIDE * 2) * 4,
              "the swizzled TMA tile must fit inside the (dead) record area");
template <int kMode>
static size_t element_smem() {
  return ((size_t)EPAIRS * XS_STRIDE + (kMode != 0 ? (size_t)EPAIRS * PAIR_STRIDE * 2 : 0)) * sizeof(float);
}

// ---- lattice fold: one thread per (row, block component) --------------------------------------------------------
// On a lattice every row has <= 27 blocks, one per lattice offset of the column vertex, and node b of a hexahedron in
// which the row vertex is node a always lands in block offset_slot(a, b).  Thread (lane = row, warp = component c) adds
// the <= 64 contributions of its component with compile-time plane offsets: a warp-wide load reads 32 consecutive
// hexahedra of one plane.  The sums are parked in shared memory at their sorted position inside the row (row record:
// lattice offset -> position among the row's L blocks, D, U blocks) and leave as the row's three contiguous segments.
constexpr int FL_ROWS = 32;
constexpr int FL_THREADS = 9 * 32;
constexpr int RR_WORDS = FEMB_ROW_REC_BYTES / 4;
struct FoldLatArgs {
  int n_vert;
  const int* inc8; const unsigned* row_rec;
  const float* em;
  const float* diag_add; float* values;
  int off_l, off_d, off_u;
  int prefetch_tiles;   // > 0: every CTA pulls the incidence lists and row records of the tile that far ahead into L2
};
template <int kMinBlocks>
__global__ void __launch_bounds__(FL_THREADS, kMinBlocks) fold_lattice_kernel(FoldLatArgs A, int v_begin, int v_end) {
  __shared__ float st[FL_ROWS * 243];
  __shared__ unsigned rr[FL_ROWS * RR_WORDS];
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int v0 = v_begin + blockIdx.x * FL_ROWS, v = v0 + lane;
  const bool live = v < v_end;
  if (A.prefetch_tiles > 0 && threadIdx.x < 20) {
    // the tile's first dependent load chain (incidence list -> element matrices) starts with a DRAM miss; a tile a quarter
    // wave ahead finds its 1 KB of incidence lists (8 lines) and 1.5 KB of row records (12 lines) in L2 instead
    const int vp = v0 + A.prefetch_tiles * FL_ROWS;
    if (vp + FL_ROWS <= v_end) {
      const char* p = threadIdx.x < 8 ? reinterpret_cast<const char*>(A.inc8 + (size_t)vp * 8) + threadIdx.x * 128
                                      : reinterpret_cast<const char*>(A.row_rec + (size_t)vp * RR_WORDS) + (threadIdx.x - 8) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
  int inc[8];                                                   // the head of the dependent chain: issued first
  {
    int4 i0 = make_int4(-1, -1, -1, -1), i1 = i0;
    if (live) {
      const int4* p = reinterpret_cast<const int4*>(A.inc8 + (size_t)v * 8);
      i0 = __ldg(p); i1 = __ldg(p + 1);
    }
    inc[0] = i0.x; inc[1] = i0.y; inc[2] = i0.z; inc[3] = i0.w; inc[4] = i1.x; inc[5] = i1.y; inc[6] = i1.z; inc[7] = i1.w;
  }
  // Row records: into registers now, into shared memory after the accumulation -- a warp issues in order, so a shared-memory
  // store right behind its load would hold back the element loads below for a memory latency.  Same for the diagonal
  // addend (pin / mass term of row r, components 0, 4, 8): its row index is requested here, the value half way through the
  // accumulation, and it is added after the last contribution (the order the output loop used to add it in).
  constexpr int RR_TILE = FL_ROWS * RR_WORDS;
  static_assert(RR_TILE <= 2 * FL_THREADS, "two row-record words per thread");
  unsigned rq0 = 0u, rq1 = 0u;
  if (v0 + (int)threadIdx.x / RR_WORDS < v_end) rq0 = __ldg(A.row_rec + (size_t)v0 * RR_WORDS + threadIdx.x);
  if (threadIdx.x + FL_THREADS < RR_TILE && v0 + (int)(threadIdx.x + FL_THREADS) / RR_WORDS < v_end)
    rq1 = __ldg(A.row_rec + (size_t)v0 * RR_WORDS + threadIdx.x + FL_THREADS);
  const bool diag_warp = A.diag_add != nullptr && (c == 0 || c == 4 || c == 8);
  int r_mine = 0;
  if (diag_warp && live) r_mine = (int)__ldg(A.row_rec + (size_t)v * RR_WORDS + 8);
  float dg = 0.f;
  const int cT = (c % 3) * 3 + c / 3;                           // the mirrored blocks are stored transposed
  // Blocked plane-major scratch: hexahedron e lives in block e >> 5 (360 planes x 32 hexahedra = 46 080 contiguous bytes) at
  // lane e & 31.  A thread's 8 loads for local node a are one base pointer + compile-time offsets (the launcher checks
  // 360 * roundup32(n_hex) < 2^32, so the element offset fits 32 bits).
  const float* __restrict__ em = A.em;
  float acc[27];
#pragma unroll
  for (int s = 0; s < 27; ++s) acc[s] = 0.f;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int e = inc[a];
    const bool have = e >= 0;
    const unsigned eo = have ? ((unsigned)e >> 5) * (unsigned)EM_BLOCK + ((unsigned)e & 31u) : 0u;
    const float* pd = em + eo + c * 32;
    const float* pm = em + eo + cT * 32;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int k = (b - a) & 7;
      const float* p = k <= 4 ? pd + ((a * 5 + k) * 9) * 32 : pm + ((b * 5 + (8 - k)) * 9) * 32;
      const float val = have ? __ldg(p) : 0.f;
      acc[offset_slot(a, b)] += val;
    }
    if (a == 3 && diag_warp && live) dg = __ldg(A.diag_add + r_mine);
  }
  if (diag_warp) acc[13] += dg;                                 // offset_slot(a, a): the diagonal block
  rr[threadIdx.x] = rq0;
  if (threadIdx.x + FL_THREADS < RR_TILE) rr[threadIdx.x + FL_THREADS] = rq1;
  __syncthreads();
  if (live) {
    const unsigned* w = rr + lane * RR_WORDS;
    unsigned pw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) pw[i] = w[i];
    float* dst = st + lane * 243 + c;
#pragma unroll
    for (int s = 0; s < 27; ++s) {
      const unsigned pos = (pw[s >> 2] >> (8 * (s & 3))) & 255u;
      if (pos < 27u) dst[pos * 9] = acc[s];
    }
  }
  __syncthreads();
  // the row's L blocks, D block and U blocks are contiguous in shared memory (<= 243 floats) and go to three places of the
  // value array: one unrolled pass with 32-bit value indices (the launcher checks 9 nnz < 2^32)
  for (int row = c; row < FL_ROWS && v0 + row < v_end; row += FL_THREADS / 32) {
    const unsigned* w = rr + row * RR_WORDS;
    const unsigned nL9 = (w[11] & 255u) * 9u, nD9 = nL9 + 9u, tot = nD9 + ((w[11] >> 8) & 255u) * 9u;
    const unsigned gL = ((unsigned)A.off_l + w[9]) * 9u, gD = ((unsigned)A.off_d + w[8]) * 9u - nL9, gU = ((unsigned)A.off_u + w[10]) * 9u - nD9;
    const float* src = st + row * 243;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const unsigned i = (unsigned)lane + 32u * it;
      if (i < tot) A.values[i < nL9 ? gL + i : (i < nD9 ? gD + i : gU + i)] = src[i];
    }
  }
}

// setup of the lattice fold: inc8[v][a] = the hexahedron in which vertex v is local node a (-1: none); row record of v:
// bytes 0..26 lattice offset slot -> sorted position inside the row (255: no such block), int32 words 7..11
// {v, index-order row, L_ptr[row], U_ptr[row], nL | nU << 8}
__global__ void __launch_bounds__(256) lat_init_kernel(int n_vert, const int* __restrict__ v2i, const int* __restrict__ L_ptr,
                                                      const int* __restrict__ U_ptr, int* __restrict__ inc8,
                                                      unsigned* __restrict__ rec, int* __restrict__ bad) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int r = __ldg(v2i + v);
  const int lp = __ldg(L_ptr + r), nl = __ldg(L_ptr + r + 1) - lp, up = __ldg(U_ptr + r), nu = __ldg(U_ptr + r + 1) - up;
  if (nl + 1 + nu > 27) *bad = 1;
  unsigned* o = rec + (size_t)v * RR_WORDS;
#pragma unroll
  for (int i = 0; i < 7; ++i) o[i] = 0xffffffffu;
  o[7] = (unsigned)v; o[8] = (unsigned)r; o[9] = (unsigned)lp; o[10] = (unsigned)up;
  o[11] = (unsigned)((nl & 255) | ((nu & 255) << 8));
#pragma unroll
  for (int a = 0; a < 8; ++a) inc8[(size_t)v * 8 + a] = -1;
}
// kCheck = false: write; true: verify that every writer agreed (a mesh that is not a lattice subset sets *bad)
template <bool kCheck>
__global__ void __launch_bounds__(256) lat_fill_kernel(int n_hex, const int* __restrict__ hex, const int* __restrict__ v2i,
                                                      const int* __restrict__ L_ptr, const int* __restrict__ U_ptr,
                                                      const int* __restrict__ elem2nnz, int off_l, int off_d, int off_u,
                                                      int* __restrict__ inc8, unsigned char* __restrict__ rec, int* __restrict__ bad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_hex * 8) return;
  const int a = (int)(t & 7), e = (int)(t >> 3);
  const int v = __ldg(hex + t);
  const int r = __ldg(v2i + v);
  const int lp = __ldg(L_ptr + r), nl = __ldg(L_ptr + r + 1) - lp, up = __ldg(U_ptr + r);
  if (kCheck) { if (inc8[(size_t)v * 8 + a] != e) *bad = 1; }
  else inc8[(size_t)v * 8 + a] = e;
  unsigned char* o = rec + (size_t)v * FEMB_ROW_REC_BYTES;
  for (int b = 0; b < 8; ++b) {
    const int s = __ldg(elem2nnz + t * 8 + b);
    const int loc = s < off_d ? (s - off_l - lp) : (s < off_u ? nl : nl + 1 + (s - off_u - up));
    if (loc < 0 || loc >= 27) { *bad = 1; continue; }
    unsigned char* d = o + offset_slot(a, b);
    if (kCheck) { if (*d != (unsigned char)loc) *bad = 1; }
    else *d = (unsigned char)loc;
  }
}

// two-kernel path: one warp per row
constexpr int AS_WARPS = 8;
constexpr int AS_THREADS = AS_WARPS * 32;
// kMinBlocks CTAs per SM: 4 -> 64 registers, no spills, 32 warps per SM; 5 -> 48 registers, a few spilled words, 40 warps
template <int kMinBlocks>
__global__ void __launch_bounds__(AS_THREADS, kMinBlocks) assemble_rows_kernel(AsmArgs A, int v_begin, int v_end, int rows_per_warp) {
  __shared__ AsmWarpSmem sm[AS_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const AsmLane LN = asm_lane(lane);
  // vertex order: the 8 rows that share a hexahedron are folded close in time, so every stored half row (read by
  // its own row and, for the mirrored blocks, by up to three other rows of the hexahedron) comes from DRAM once.
  // A warp folds `rows_per_warp` consecutive vertices; the plan record of the next one is fetched while the current
  // one is folded, which leaves one exposed memory round trip per row (the element matrices) instead of two.
  const int v0 = v_begin + (blockIdx.x * AS_WARPS + w) * rows_per_warp;
  const int v1 = min(v0 + rows_per_warp, v_end);
  if (v0 >= v1) return;
  AsmMeta M = asm_meta(A, v0, lane);
  if (lane < 12) sm[w].stage[AS_ZERO + lane] = 0.f;
  for (int v = v0; v < v1; ++v) {
    AsmMeta Mn = M;
    if (v + 1 < v1) Mn = asm_meta(A, v + 1, lane);
    assemble_row(A, LN, M, lane, sm[w]);
    M = Mn;
  }
}

// out[row] = sum of the element forces of the incident hexahedra + gravity + pin spring
__global__ void __launch_bounds__(256) force_gather_kernel(int n_vert, const int* __restrict__ vert_elem,
                                                          const float* __restrict__ elem_f, const float* __restrict__ x,
                                                          const float* __restrict__ mass, float g,
                                                          const int* __restrict__ pin, const float* __restrict__ pin_pos,
                                                          float control_mag, const int* __restrict__ v2i,
                                                          float* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  float f[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ea = __ldg(vert_elem + (size_t)v * 8 + j);
    if (ea >= 0) {
      const float* s = elem_f + (size_t)ea * 3;
      f[0] += __ldg(s); f[1] += __ldg(s + 1); f[2] += __ldg(s + 2);
    }
  }
  if (mass != nullptr) f[1] += __ldg(mass + v) * g;                                  // quai_hexagon.py:176-180
  if (pin != nullptr && __ldg(pin + v) == 1) {                                       // quai_hexagon.py:160-165
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] += control_mag * (__ldg(pin_pos + 3 * (size_t)v + c) - __ldg(x + 3 * (size_t)v + c));
  }
  const int row = v2i != nullptr ? __ldg(v2i + v) : v;
  out[3 * (size_t)row] = f[0]; out[3 * (size_t)row + 1] = f[1]; out[3 * (size_t)row + 2] = f[2];
}

// rest Jacobian per (hexahedron, Gauss point): det and inverse                (quai_hexagon.py:4-18)
__global__ void __launch_bounds__(256) prepare_rest_kernel(int n_hex, const int* __restrict__ hex,
                                                          const float* __restrict__ x, float* __restrict__ detJ,
                                                          float* __restrict__ invJ) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_hex * 8) return;
  const int e = (int)(t >> 3), q = (int)(t & 7);
  float J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = 0; a < 8; ++a) {
    const int v = __ldg(hex + (size_t)e * 8 + a);
    float dN[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = ((a >> (2 - c)) & 1) ? 1.0f : -1.0f;
      const int d0 = (c == 0) ? 1 : 0, d1 = (c == 2) ? 1 : 2;
      const float f0 = (((a >> (2 - d0)) & 1) == ((q >> (2 - d0)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
      const float f1 = (((a >> (2 - d1)) & 1) == ((q >> (2 - d1)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
      dN[c] = (s * f0) * f1 * 0.125f;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float xr = __ldg(x + 3 * (size_t)v + r);
#pragma unroll
      for (int c = 0; c < 3; ++c) J[3 * r + c] += xr * dN[c];
    }
  }
  const float c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  const float det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  const float id = 1.0f / det;
  float* o = invJ + (size_t)t * 9;
  o[0] = c00 * id; o[1] = (J[2] * J[7] - J[1] * J[8]) * id; o[2] = (J[1] * J[5] - J[2] * J[4]) * id;
  o[3] = c01 * id; o[4] = (J[0] * J[8] - J[2] * J[6]) * id; o[5] = (J[2] * J[3] - J[0] * J[5]) * id;
  o[6] = c02 * id; o[7] = (J[1] * J[6] - J[0] * J[7]) * id; o[8] = (J[0] * J[4] - J[1] * J[3]) * id;
  detJ[t] = det;
}

// Sum_8 + prepare_mass as gathers                       (cublas.py:215-218, quai_hexagon.py:20-26)
__global__ void volume_kernel(int n_hex, const float* __restrict__ detJ, float* __restrict__ vol) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_hex) return;
  const float* d = detJ + (size_t)e * 8;
  vol[e] = d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7];
}
__global__ void mass_kernel(int n_vert, const int* __restrict__ vert_elem, const float* __restrict__ vol,
                            float* __restrict__ mass) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  float m = 0.f;
  for (int j = 0; j < 8; ++j) {
    const int ea = __ldg(vert_elem + (size_t)v * 8 + j);
    if (ea >= 0) m += __ldg(vol + (ea >> 3));
  }
  mass[v] = m;
}

// energy: partial sums per block, then one block adds them in a fixed order
constexpr int EN_HEX = ETHREADS / 8;                  // hexahedra per group of the energy kernel (8 lanes each, scalar arithmetic)
__global__ void __launch_bounds__(ETHREADS) elastic_energy_kernel(ElemIn in, int n_groups, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const int sub = threadIdx.x & 7;
  float dN[8][3];
  shape_grads(sub, dN);
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int e = grp * EN_HEX + (threadIdx.x >> 3);
    const bool live = e < in.n_hex;
    const int ee = live ? e : in.n_hex - 1;
    const int vtx = __ldg(in.hex + (size_t)ee * 8 + sub);
    float xn[3] = {__ldg(in.x + 3 * (size_t)vtx), __ldg(in.x + 3 * (size_t)vtx + 1), __ldg(in.x + 3 * (size_t)vtx + 2)};
    float X[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) X[a][c] = __shfl_sync(0xffffffffu, xn[c], a, 8);
    float iJ[9];
    const float* ip = in.invJ + ((size_t)ee * 8 + sub) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) iJ[k] = __ldg(ip + k);
    const float det = __ldg(in.detJ + (size_t)ee * 8 + sub);
    float fq[8][3];
    float en;
    gauss_point<float, false, false, true>(X, dN, iJ, det, in.mu, in.la, nullptr, fq, en);
    if (live) acc += en;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) vertex_energy_kernel(int n_vert, const float* __restrict__ x,
                                                           const float* __restrict__ mass, float g,
                                                           const int* __restrict__ pin, const float* __restrict__ pin_pos,
                                                           float control_mag, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vert; v += gridDim.x * blockDim.x) {
    float en = 0.f;
    if (mass != nullptr) en = -__ldg(mass + v) * g * __ldg(x + 3 * (size_t)v + 1);                 // quai_hexagon.py:60
    if (pin != nullptr && __ldg(pin + v) == 1) {
      float d2 = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = __ldg(pin_pos + 3 * (size_t)v + c) - __ldg(x + 3 * (size_t)v + c);
        d2 += d * d;
      }
      en += 0.5f * control_mag * d2;
    }
    acc += en;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) final_sum_kernel(int n, const float* __restrict__ partial, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256) coarse_pin_kernel(int n_coarse, const int* __restrict__ off, const int* __restrict__ col,
                                                        const float* __restrict__ w, const int* __restrict__ pin_fine,
                                                        const int* __restrict__ v2i_c, int off_d, float control_mag,
                                                        float* __restrict__ values) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  if (J >= n_coarse) return;
  float s = 0.f;
  for (int p = __ldg(off + J); p < __ldg(off + J + 1); ++p)
    if (__ldg(pin_fine + __ldg(col + p)) == 1) {
      const float ww = __ldg(w + p);
      s += ww * ww * control_mag;                                                   // cublas.py:242
    }
  if (s != 0.f) {
    float* d = values + ((size_t)off_d + __ldg(v2i_c + J)) * 9;
    d[0] += s; d[4] += s; d[8] += s;
  }
}

static size_t em_stride_of(int n_hex) { return ((size_t)n_hex + 31) / 32 * 32; }
// (the lattice fold addresses the plane-major scratch with 32-bit element offsets)
static bool lattice_available(const FembLevel* l) {
  // (32-bit element offsets into the scratch and 32-bit indices into the value array)
  return l->lat_inc8 != nullptr && l->lat_row_rec != nullptr && (unsigned long long)EM_PLANES * em_stride_of(l->n_hex) < (1ull << 32) &&
         (unsigned long long)l->nnz * 9ull < (1ull << 32);
}

static ElemIn make_in(const FembLevel* l, const float* x, float mu, float la) {
  ElemIn in;
  in.hex = l->hex; in.x = x; in.invJ = l->invJ; in.detJ = l->detJ; in.e_begin = 0; in.n_hex = l->n_hex; in.mu = mu; in.la = la;
  in.prefetch_ctas = 0;
  return in;
}

static AsmArgs make_asm(const FembLevel* l, const float* diag_add, const float* elem_k, float* values) {
  AsmArgs A;
  A.n_vert = l->n_vert; A.i2v = l->i2v; A.v2i = l->v2i; A.vert_elem = l->vert_elem; A.elem2nnz = l->elem2nnz;
  A.L_ptr = l->L_ptr; A.U_ptr = l->U_ptr; A.off_l = l->off_l; A.off_d = l->off_d; A.off_u = l->off_u;
  A.fold = l->fold_plan; A.elem_k = elem_k; A.diag_add = diag_add; A.values = values;
  return A;
}

// ---- tensor maps over the plane-major scratch (host side) ------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so that libcuda is not a link
// dependency (the library is built on machines without a driver).  No encoder -> the non-TMA kernels run.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
static int g_asm_async = 3;   // 0: store loop, 1: TMA tensor store, 2: + rest data through bulk copies, 3: + L2 prefetch of index data
struct EmMaps {
  CUtensorMap store;    // 3-D view {hexahedron in block, plane, block} of the blocked scratch, box {32, 120, 1}, 128-byte swizzle:
                        // a third of the element kernel's tile
  bool ok;
};
static EmMaps make_em_maps(const float* elem_k, int n_hex) {
  EmMaps M;
  memset(&M, 0, sizeof(M));
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || elem_k == nullptr || (reinterpret_cast<uintptr_t>(elem_k) & 15u) != 0 || n_hex <= 0) return M;
  void* base = const_cast<float*>(elem_k);
  const cuuint32_t one3[3] = {1, 1, 1};
  const cuuint64_t d3[3] = {32, (cuuint64_t)EM_PLANES, (cuuint64_t)(em_stride_of(n_hex) / 32)};
  const cuuint64_t s3[2] = {32 * sizeof(float), EM_BLOCK * sizeof(float)};
  const cuuint32_t b3[3] = {32, (cuuint32_t)EM_TMA_BOX_ROWS, 1};
  M.ok = enc(&M.store, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, d3, s3, b3, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  return M;
}

template <int kMode, bool kSoA, bool kLa, bool kTma, bool kBulk>
static int launch_element_la(ElemIn in, int e_begin, int e_end, float* elem_f, float* elem_k, const CUtensorMap& map, cudaStream_t st) {
  static bool configured_dev[64] = {};                         // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  const size_t smem = element_smem<kMode>();
  if (!configured) {
    cudaFuncSetAttribute(element_kernel<kMode, kSoA, kLa, kTma, kBulk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  if (e_end <= e_begin) return FEMB_OK;
  in.e_begin = e_begin; in.n_hex = e_end;
  element_kernel<kMode, kSoA, kLa, kTma, kBulk><<<cdiv(e_end - e_begin, EPB), ETHREADS, smem, st>>>(in, elem_f, elem_k, map);
  count_launch();
  return check_launch("element_kernel");
}
// maps == nullptr (or not usable): the tile leaves through the store loop and the rest data comes through per-lane loads
template <int kMode, bool kSoA>
static int launch_element(ElemIn in, int e_begin, int e_end, float* elem_f, float* elem_k, cudaStream_t st, const EmMaps* maps = nullptr) {
  static const CUtensorMap no_map = {};
  const bool la = kMode != 0 && in.la != 0.f;
  constexpr bool kCanTma = kSoA && kMode != 0;
  in.prefetch_ctas = g_asm_async >= 3 ? 111 : 0;            // a quarter of the 148 x 3 resident CTAs
  if (kCanTma && maps != nullptr && maps->ok && g_asm_async) {
    const bool bulk = g_asm_async >= 2 && ((reinterpret_cast<uintptr_t>(in.invJ) | reinterpret_cast<uintptr_t>(in.detJ)) & 15u) == 0;
    if (bulk) {
      if (la) return launch_element_la<kMode, kSoA, true, kCanTma, kCanTma>(in, e_begin, e_end, elem_f, elem_k, maps->store, st);
      return launch_element_la<kMode, kSoA, false, kCanTma, kCanTma>(in, e_begin, e_end, elem_f, elem_k, maps->store, st);
    }
    if (la) return launch_element_la<kMode, kSoA, true, kCanTma, false>(in, e_begin, e_end, elem_f, elem_k, maps->store, st);
    return launch_element_la<kMode, kSoA, false, kCanTma, false>(in, e_begin, e_end, elem_f, elem_k, maps->store, st);
  }
  if (la) return launch_element_la<kMode, kSoA, true, false, false>(in, e_begin, e_end, elem_f, elem_k, no_map, st);
  return launch_element_la<kMode, kSoA, false, false, false>(in, e_begin, e_end, elem_f, elem_k, no_map, st);
}

static int launch_force_gather(const FembLevel* l, const float* x, const float* mass, float g, const int* pin,
                               const float* pin_pos, float control_mag, int ordered, const float* elem_f, float* out,
                               cudaStream_t st) {
  force_gather_kernel<<<cdiv(l->n_vert, 256), 256, 0, st>>>(l->n_vert, l->vert_elem, elem_f, x, mass, g,
                                                             pin_pos ? pin : nullptr, pin_pos, control_mag,
                                                             ordered ? l->v2i : nullptr, out);
  count_launch();
  return check_launch("force_gather_kernel");
}

static int g_fold_rows_per_warp = 4, g_fold_min_blocks = 4;
static int launch_assemble(const AsmArgs& as, int v_begin, int v_end, cudaStream_t st) {
  if (v_end <= v_begin) return FEMB_OK;
  const int rpw = g_fold_rows_per_warp;
  const unsigned grid = cdiv(v_end - v_begin, AS_WARPS * rpw);
  if (g_fold_min_blocks == 5) assemble_rows_kernel<5><<<grid, AS_THREADS, 0, st>>>(as, v_begin, v_end, rpw);
  else assemble_rows_kernel<4><<<grid, AS_THREADS, 0, st>>>(as, v_begin, v_end, rpw);
  count_launch();
  return check_launch("assemble_rows_kernel");
}

static int g_lat_min_blocks = 4;
static int launch_fold_lattice(const FembLevel* l, const float* diag_add, const float* elem_k, float* values, cudaStream_t st) {
  FoldLatArgs A;
  A.n_vert = l->n_vert; A.inc8 = l->lat_inc8; A.row_rec = reinterpret_cast<const unsigned*>(l->lat_row_rec);
  A.em = elem_k; A.diag_add = diag_add; A.values = values;
  A.off_l = l->off_l; A.off_d = l->off_d; A.off_u = l->off_u; A.prefetch_tiles = g_asm_async >= 3 ? 148 : 0;   // a quarter of the 148 x 4 resident CTAs
  const unsigned grid = cdiv(l->n_vert, FL_ROWS);
  if (g_lat_min_blocks == 3) fold_lattice_kernel<3><<<grid, FL_THREADS, 0, st>>>(A, 0, l->n_vert);
  else if (g_lat_min_blocks == 5) fold_lattice_kernel<5><<<grid, FL_THREADS, 0, st>>>(A, 0, l->n_vert);
  else fold_lattice_kernel<4><<<grid, FL_THREADS, 0, st>>>(A, 0, l->n_vert);
  count_launch();
  return check_launch("fold_lattice_kernel");
}

// element matrices of all hexahedra + fold of all rows; kMode 1 = Hessian only, 2 = element forces as well
template <int kMode>
static int hessian_pipeline(const FembLevel* l, const float* x, float mu, float la, const float* diag_add, float* elem_f,
                            float* elem_k, float* values, cudaStream_t st) {
  const ElemIn in = make_in(l, x, mu, la);
  // (A chunked two-stream software pipeline of the two stages and a one-kernel variant with the fold inside the SM were
  // built and measured slower)
  if (lattice_available(l)) {
    const EmMaps maps = g_asm_async ? make_em_maps(elem_k, l->n_hex) : EmMaps{};
    int rc = launch_element<kMode, true>(in, 0, l->n_hex, elem_f, elem_k, st, &maps);
    if (rc) return rc;
    return launch_fold_lattice(l, diag_add, elem_k, values, st);
  }
  if (!(l->elem2nnz && l->vert_elem && l->i2v && l->fold_plan)) {
    set_error("assembly: level has neither a lattice fold nor a fold plan");
    return FEMB_E_INVALID;
  }
  const AsmArgs as = make_asm(l, diag_add, elem_k, values);
  int rc = launch_element<kMode, false>(in, 0, l->n_hex, elem_f, elem_k, st);
  if (rc) return rc;
  return launch_assemble(as, 0, l->n_vert, st);
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_set_fold_config(int rows_per_warp, int min_blocks_per_sm) {
  if (rows_per_warp >= 1 && rows_per_warp <= 64) g_fold_rows_per_warp = rows_per_warp;
  if (min_blocks_per_sm == 4 || min_blocks_per_sm == 5) g_fold_min_blocks = min_blocks_per_sm;
  if (min_blocks_per_sm >= 3 && min_blocks_per_sm <= 5) g_lat_min_blocks = min_blocks_per_sm;
  return FEMB_OK;
}

int femb_set_assembly_async(int mode) {
  if (mode >= 0 && mode <= 3) g_asm_async = mode;
  return FEMB_OK;
}

int femb_prepare_rest(int n_hex, const int* hex, const float* x_rest, float* detJ, float* invJ, void* stream) {
  FEMB_REQUIRE(n_hex >= 0 && hex && x_rest && detJ && invJ, "null pointer or negative size");
  if (n_hex == 0) return FEMB_OK;
  prepare_rest_kernel<<<cdiv((long long)n_hex * 8, 256), 256, 0, as_stream(stream)>>>(n_hex, hex, x_rest, detJ, invJ);
  count_launch();
  return check_launch("prepare_rest_kernel");
}

int femb_build_fold_plan(const FembLevel* lvl, unsigned char* fold_plan, void* stream) {
  FEMB_REQUIRE(lvl && fold_plan, "null pointer");
  FEMB_REQUIRE(lvl->v2i && lvl->vert_elem && lvl->elem2nnz && lvl->L_ptr && lvl->U_ptr, "level not initialised");
  if (lvl->n_vert == 0) return FEMB_OK;
  fold_plan_kernel<<<cdiv(lvl->n_vert, 8), 256, 0, as_stream(stream)>>>(make_asm(lvl, nullptr, nullptr, nullptr), fold_plan);
  count_launch();
  return check_launch("fold_plan_kernel");
}

int femb_lumped_mass(const FembLevel* lvl, float* vol, float* mass, void* stream) {
  FEMB_REQUIRE(lvl && vol && mass && lvl->detJ && lvl->vert_elem, "null pointer");
  if (lvl->n_hex > 0) {
    volume_kernel<<<cdiv(lvl->n_hex, 256), 256, 0, as_stream(stream)>>>(lvl->n_hex, lvl->detJ, vol);
    count_launch();
  }
  if (lvl->n_vert > 0) {
    mass_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->vert_elem, vol, mass);
    count_launch();
  }
  return check_launch("lumped_mass");
}

int femb_force(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g, const int* pin,
               const float* pin_pos, float control_mag, int ordered, float* elem_f, float* out, void* stream) {
  FEMB_REQUIRE(lvl && x && elem_f && out, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->vert_elem && lvl->v2i, "level not initialised");
  if (lvl->n_hex > 0) {
    int rc = launch_element<0, false>(make_in(lvl, x, mu, lambda), 0, lvl->n_hex, elem_f, nullptr, as_stream(stream));
    if (rc) return rc;
  }
  if (lvl->n_vert == 0) return FEMB_OK;
  return launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, ordered, elem_f, out, as_stream(stream));
}

int femb_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* diag_add, float* elem_k,
                 float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && values && elem_k, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->v2i, "level not initialised");
  return hessian_pipeline<1>(lvl, x, mu, lambda, diag_add, nullptr, elem_k, values, as_stream(stream));
}

int femb_force_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
                       const int* pin, const float* pin_pos, float control_mag, const float* diag_add, float* elem_f,
                       float* elem_k, float* force_out, float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && force_out && values && elem_f && elem_k, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->vert_elem && lvl->v2i, "level not initialised");
  int rc = hessian_pipeline<2>(lvl, x, mu, lambda, diag_add, elem_f, elem_k, values, as_stream(stream));
  if (rc) return rc;
  return launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, 1, elem_f, force_out, as_stream(stream));
}

int femb_build_lattice_fold(const FembLevel* lvl, int* inc8, unsigned char* row_rec, int* bad, void* stream) {
  FEMB_REQUIRE(lvl && inc8 && row_rec && bad, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->v2i && lvl->elem2nnz && lvl->L_ptr && lvl->U_ptr, "level not initialised");
  cudaStream_t st = as_stream(stream);
  if (lvl->n_vert > 0) {
    lat_init_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, st>>>(lvl->n_vert, lvl->v2i, lvl->L_ptr, lvl->U_ptr, inc8,
                                                           reinterpret_cast<unsigned*>(row_rec), bad);
    count_launch();
  }
  if (lvl->n_hex > 0) {
    const unsigned grid = cdiv((long long)lvl->n_hex * 8, 256);
    lat_fill_kernel<false><<<grid, 256, 0, st>>>(lvl->n_hex, lvl->hex, lvl->v2i, lvl->L_ptr, lvl->U_ptr, lvl->elem2nnz, lvl->off_l,
                                                 lvl->off_d, lvl->off_u, inc8, row_rec, bad);
    lat_fill_kernel<true><<<grid, 256, 0, st>>>(lvl->n_hex, lvl->hex, lvl->v2i, lvl->L_ptr, lvl->U_ptr, lvl->elem2nnz, lvl->off_l,
                                                lvl->off_d, lvl->off_u, inc8, row_rec, bad);
    count_launch(2);
  }
  return check_launch("femb_build_lattice_fold");
}

int femb_energy(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g, const int* pin,
                const float* pin_pos, float control_mag, float* scratch, float* energy, void* stream) {
  FEMB_REQUIRE(lvl && x && scratch && energy, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int n_groups = (int)cdiv(lvl->n_hex, EN_HEX);
  const int gb = n_groups < 512 ? (n_groups > 0 ? n_groups : 1) : 512;
  elastic_energy_kernel<<<gb, ETHREADS, 0, st>>>(make_in(lvl, x, mu, lambda), n_groups, scratch);
  int vb = (int)cdiv(lvl->n_vert, 256);
  vb = vb < 512 ? (vb > 0 ? vb : 1) : 512;
  vertex_energy_kernel<<<vb, 256, 0, st>>>(lvl->n_vert, x, mass, g, pin_pos ? pin : nullptr, pin_pos, control_mag,
                                           scratch + gb);
  final_sum_kernel<<<1, 256, 0, st>>>(gb + vb, scratch, energy);
  count_launch(3);
  return check_launch("energy");
}

int femb_coarse_pin_penalty(const FembLevel* coarse, const FembTransfer* t, const int* pin_fine, float control_mag,
                            float* values, void* stream) {
  FEMB_REQUIRE(coarse && t && pin_fine && values, "null pointer");
  FEMB_REQUIRE(t->nrow == coarse->n_vert, "transfer rows != coarse vertices");
  if (t->nrow == 0) return FEMB_OK;
  coarse_pin_kernel<<<cdiv(t->nrow, 256), 256, 0, as_stream(stream)>>>(t->nrow, t->offsets, t->columns, t->weights, pin_fine,
                                                                        coarse->v2i, coarse->off_d, control_mag, values);
  count_launch();
  return check_launch("coarse_pin_kernel");
}

}  // extern "C"
