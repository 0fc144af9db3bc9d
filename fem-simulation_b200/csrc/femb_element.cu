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
// B200 design: 8 lanes per hexahedron (4 hexahedra per warp).  In phase A lane q evaluates everything
// that depends only on the Gauss point q (F, S, F F^T, g_b = J^-T dN_b, u_b = F g_b, t_b = S g_b) once
// and parks it in shared memory; in phase B lane a accumulates, over the 8 Gauss points, the 5 blocks
// K[a][(a+k) & 7], k = 0..4 (the element matrix is symmetric: the cyclic half is enough) with the closed form
//     K_ab[j][i] = det * ( (g_a . S g_b) d_ji + mu (g_a . g_b)(F F^T)_ji + mu u_b[j] u_a[i] + la u_a[j] u_b[i] )
// (== d(P g_a det)_j / d x_b,i of quai_hexagon.py:86-106).  The half rows go to a scratch array (192 bytes per
// row, written as coalesced 512-byte stores) and are summed into the fixed [L;D;U] value layout by a per-row
// gather over the fold plan, the precomputed inverse of the element->nnz map (fixed order, no atomics).
#include <cstdio>

#include "femb_common.cuh"

namespace femb {

constexpr int EPB = FEMB_ASM_CHUNK_HEXES;            // hexahedra per CTA
constexpr int ETHREADS = EPB * 8;  // 128 threads

// per (hexahedron, Gauss point) record in shared memory, in floats; written and read as float4:
//   [0..5] det*mu * F F^T (symmetric: 00 01 02 11 12 22)   [6] det*mu   [7] det*la
//   [8 + 6 b ..] per node b: g_b[3], u_b[3]          (48 floats; broadcast-read by all 8 lanes)
//   [56 + 3 b ..] per node b: det * S g_b[3]          (24 floats; lane a reads only its own t_a)
constexpr int QP_FFT = 0;
constexpr int QP_C = 6;
constexpr int QP_GU = 8;
constexpr int QP_T = 56;
constexpr int QP_SIZE = 84;   // 21 float4: 16-byte accesses of 8 consecutive lanes hit distinct banks (21 is odd)
// floats between the records of consecutive hexahedra: 8*84 + 16 = 16 (mod 32), so the per-lane 8-byte reads of phase B
// (two hexahedra per half-warp, 16 banks each) are conflict-free
constexpr int HEX_STRIDE = 8 * QP_SIZE + 16;
constexpr int EK_ROW = FEMB_ELEM_K_FLOATS / 8;   // floats per stored row of an element matrix (5 blocks + 3 pad)

struct ElemIn {
  const int* hex;
  const float* x;
  const float* invJ;
  const float* detJ;
  int e_begin, n_hex;    // the launch covers hexahedra [e_begin, n_hex)
  float mu, la;
};

// Phase A for lane q of hexahedron e.  Writes the Gauss-point record and returns P*det (first
// Piola-Kirchhoff stress times det J) plus the energy density times det.
__device__ __forceinline__ void gauss_point(const ElemIn& in, int e, int q, const float (&X)[8][3], float* rec,
                                            float (&fq)[8][3], float& energy, bool want_rec) {
  // shape-function gradients at this Gauss point (q is a runtime value here)
  float dN[8][3];
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
  float iJ[9];
  const float* ip = in.invJ + ((size_t)e * 8 + q) * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) iJ[k] = __ldg(ip + k);
  const float det = __ldg(in.detJ + (size_t)e * 8 + q);
  // F0 = sum_a x_a (x) dN_a ; F = F0 * invJ                                  (quai_hexagon.py:77-84)
  float F0[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = 0.f;
#pragma unroll
      for (int a = 0; a < 8; ++a) v += X[a][r] * dN[a][c];
      F0[3 * r + c] = v;
    }
  float F[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) F[3 * r + c] = F0[3 * r] * iJ[c] + F0[3 * r + 1] * iJ[3 + c] + F0[3 * r + 2] * iJ[6 + c];
  // E = 0.5 (F^T F - I), S = 2 mu E + la tr(E) I   (symmetric: 00 01 02 11 12 22)
  float E[6];
  E[0] = 0.5f * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6] - 1.f);
  E[1] = 0.5f * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
  E[2] = 0.5f * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]);
  E[3] = 0.5f * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7] - 1.f);
  E[4] = 0.5f * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
  E[5] = 0.5f * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8] - 1.f);
  const float trE = E[0] + E[3] + E[5];
  const float sumE2 = E[0] * E[0] + E[3] * E[3] + E[5] * E[5] + 2.f * (E[1] * E[1] + E[2] * E[2] + E[4] * E[4]);
  energy = (sumE2 * in.mu + 0.5f * in.la * trE * trE) * det;                 // quai_hexagon.py:47-53
  float S[6];
  const float tm = 2.f * in.mu;
  S[0] = tm * E[0] + in.la * trE; S[1] = tm * E[1]; S[2] = tm * E[2];
  S[3] = tm * E[3] + in.la * trE; S[4] = tm * E[4]; S[5] = tm * E[5] + in.la * trE;
  // P = F S
  float P[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    P[3 * r + 0] = F[3 * r] * S[0] + F[3 * r + 1] * S[1] + F[3 * r + 2] * S[2];
    P[3 * r + 1] = F[3 * r] * S[1] + F[3 * r + 1] * S[3] + F[3 * r + 2] * S[4];
    P[3 * r + 2] = F[3 * r] * S[2] + F[3 * r + 1] * S[4] + F[3 * r + 2] * S[5];
  }
  float R[QP_SIZE];
  if (want_rec) {
    const float cmu = det * in.mu;
    R[QP_FFT + 0] = cmu * (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
    R[QP_FFT + 1] = cmu * (F[0] * F[3] + F[1] * F[4] + F[2] * F[5]);
    R[QP_FFT + 2] = cmu * (F[0] * F[6] + F[1] * F[7] + F[2] * F[8]);
    R[QP_FFT + 3] = cmu * (F[3] * F[3] + F[4] * F[4] + F[5] * F[5]);
    R[QP_FFT + 4] = cmu * (F[3] * F[6] + F[4] * F[7] + F[5] * F[8]);
    R[QP_FFT + 5] = cmu * (F[6] * F[6] + F[7] * F[7] + F[8] * F[8]);
    R[QP_C + 0] = cmu;
    R[QP_C + 1] = det * in.la;
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    // g_b = invJ^T dN_b  (dF row of quai_hexagon.py:95)
    float g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = dN[b][0] * iJ[j] + dN[b][1] * iJ[3 + j] + dN[b][2] * iJ[6 + j];
    // nodal force contribution  -(P g_b) det                                 (quai_hexagon.py:128-133)
#pragma unroll
    for (int r = 0; r < 3; ++r) fq[b][r] = -det * (P[3 * r] * g[0] + P[3 * r + 1] * g[1] + P[3 * r + 2] * g[2]);
    if (want_rec) {
      float* o = R + QP_GU + 6 * b;
      o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
#pragma unroll
      for (int r = 0; r < 3; ++r) o[3 + r] = F[3 * r] * g[0] + F[3 * r + 1] * g[1] + F[3 * r + 2] * g[2];
      float* t = R + QP_T + 3 * b;
      t[0] = det * (S[0] * g[0] + S[1] * g[1] + S[2] * g[2]);
      t[1] = det * (S[1] * g[0] + S[3] * g[1] + S[4] * g[2]);
      t[2] = det * (S[2] * g[0] + S[4] * g[1] + S[5] * g[2]);
    }
  }
  if (want_rec) {
    float4* r4 = reinterpret_cast<float4*>(rec);
#pragma unroll
    for (int v = 0; v < 20; ++v) r4[v] = make_float4(R[4 * v], R[4 * v + 1], R[4 * v + 2], R[4 * v + 3]);
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

// kMode: 0 = forces only, 1 = Hessian only, 2 = both
template <int kMode>
__global__ void __launch_bounds__(ETHREADS, 5) element_kernel(ElemIn in, float* __restrict__ elem_f, float* __restrict__ elem_k) {
  constexpr bool kForce = (kMode != 1);
  constexpr bool kHess = (kMode != 0);
  extern __shared__ __align__(16) float smem[];
  float* s_rec = smem;                                                // [EPB][HEX_STRIDE]: 8 records per hexahedron (Hessian only)
  const int sub = threadIdx.x & 7;
  const int le = threadIdx.x >> 3;
  const int e = in.e_begin + blockIdx.x * EPB + le;                   // this launch covers hexahedra [e_begin, n_hex)
  const bool live = e < in.n_hex;
  const int ee = live ? e : in.n_hex - 1;                             // dead lanes shadow a valid hexahedron
  // node positions: lane `sub` loads node `sub`, the group shares them with width-8 shuffles
  const int vtx = __ldg(in.hex + (size_t)ee * 8 + sub);
  float xn[3] = {__ldg(in.x + 3 * (size_t)vtx), __ldg(in.x + 3 * (size_t)vtx + 1), __ldg(in.x + 3 * (size_t)vtx + 2)};
  float X[8][3];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) X[a][c] = __shfl_sync(0xffffffffu, xn[c], a, 8);

  {
    float fq[8][3];
    float energy;
    float* rec = s_rec + (size_t)le * HEX_STRIDE + sub * QP_SIZE;
    gauss_point(in, ee, sub, X, rec, fq, energy, kHess);
    if (kForce) {
      // lane q holds the Gauss-point contributions to all 8 nodes; a reduce-scatter over the 8 lanes of the
      // hexahedron (xor 4, 2, 1) leaves lane a with the force on node a
      float h4[4][3], h2[2][3], f[3];
      const bool up4 = (sub & 4) != 0, up2 = (sub & 2) != 0, up1 = (sub & 1) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float give = up4 ? fq[i][c] : fq[i + 4][c], keep = up4 ? fq[i + 4][c] : fq[i][c];
          h4[i][c] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
        }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float give = up2 ? h4[i][c] : h4[i + 2][c], keep = up2 ? h4[i + 2][c] : h4[i][c];
          h2[i][c] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float give = up1 ? h2[0][c] : h2[1][c], keep = up1 ? h2[1][c] : h2[0][c];
        f[c] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
      }
      if (live) {
        float* dst = elem_f + ((size_t)e * 8 + sub) * 3;
        dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2];
      }
    }
  }
  if (!kHess) return;
  __syncthreads();

  // ---- phase B: lane a accumulates the 5 blocks K[a][(a+k) & 7], k = 0..4 ------------------------------
  // K_ab[j][i] = sum_q  d_ji (t_a . g_b) + (det mu FF^T)_ji (g_a . g_b) + det mu u_b[j] u_a[i] + det la u_a[j] u_b[i]
  // with t_a = det S g_a (S symmetric: g_a . S g_b == S g_a . g_b); every term is one FFMA into an accumulator.
  // The element matrix is symmetric (K_ba = K_ab^T), so the cyclic half {(a, a+k)} is enough: the blocks
  // K[a][(a-m) & 7], m = 1..3, are the transposes of block m of lane (a-m) & 7 and arrive by shuffle.
  float K[5][9], Ks[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    Ks[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) K[k][i] = 0.f;
  }
  const bool has_la = in.la != 0.f;
#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const float* r = s_rec + (size_t)le * HEX_STRIDE + q * QP_SIZE;
    const float4* r4 = reinterpret_cast<const float4*>(r);
    const float4 h0 = r4[0], h1 = r4[1];                   // broadcast reads (all 8 lanes of a hexahedron, same address)
    const float FFT[6] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y};
    const float cmu = h1.z, cla = h1.w;
    const float2* ga2 = reinterpret_cast<const float2*>(r + QP_GU + 6 * sub);   // own node
    const float2 o0 = ga2[0], o1 = ga2[1], o2 = ga2[2];
    const float g_a[3] = {o0.x, o0.y, o1.x};
    const float ua_mu[3] = {cmu * o1.y, cmu * o2.x, cmu * o2.y};
    const float ua_la[3] = {cla * o1.y, cla * o2.x, cla * o2.y};
    const float* tp = r + QP_T + 3 * sub;
    const float t_a[3] = {tp[0], tp[1], tp[2]};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float2* gb2 = reinterpret_cast<const float2*>(r + QP_GU + 6 * ((sub + k) & 7));
      const float2 p0 = gb2[0], p1 = gb2[1], p2 = gb2[2];
      const float gb[3] = {p0.x, p0.y, p1.x};
      const float u_b[3] = {p1.y, p2.x, p2.y};
      Ks[k] = fmaf(t_a[0], gb[0], fmaf(t_a[1], gb[1], fmaf(t_a[2], gb[2], Ks[k])));
      const float s2 = fmaf(g_a[0], gb[0], fmaf(g_a[1], gb[1], g_a[2] * gb[2]));
      K[k][0] = fmaf(FFT[0], s2, fmaf(u_b[0], ua_mu[0], K[k][0]));
      K[k][1] = fmaf(FFT[1], s2, fmaf(u_b[0], ua_mu[1], K[k][1]));
      K[k][2] = fmaf(FFT[2], s2, fmaf(u_b[0], ua_mu[2], K[k][2]));
      K[k][3] = fmaf(FFT[1], s2, fmaf(u_b[1], ua_mu[0], K[k][3]));
      K[k][4] = fmaf(FFT[3], s2, fmaf(u_b[1], ua_mu[1], K[k][4]));
      K[k][5] = fmaf(FFT[4], s2, fmaf(u_b[1], ua_mu[2], K[k][5]));
      K[k][6] = fmaf(FFT[2], s2, fmaf(u_b[2], ua_mu[0], K[k][6]));
      K[k][7] = fmaf(FFT[4], s2, fmaf(u_b[2], ua_mu[1], K[k][7]));
      K[k][8] = fmaf(FFT[5], s2, fmaf(u_b[2], ua_mu[2], K[k][8]));
      if (has_la) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) K[k][3 * j + i] = fmaf(ua_la[j], u_b[i], K[k][3 * j + i]);
      }
    }
  }
  // Row a of the element matrix is stored ROTATED and HALF: position p = 0..4 holds the block of column (a + p) & 7
  // (45 floats + 3 pad = 192 bytes).  The blocks of columns (a - m) & 7, m = 1..3, are the transposes of position m of
  // row (a - m) & 7 and are picked up from there by the assembly.
  // Layout of the 48 floats: elements 0..7 of position p at [8p, 8p+8), the ninth elements at [40, 45), 3 pad.
  float row[48];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    K[k][0] += Ks[k]; K[k][4] += Ks[k]; K[k][8] += Ks[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) row[8 * k + i] = K[k][i];
    row[40 + k] = K[k][8];
  }
  row[45] = 0.f; row[46] = 0.f; row[47] = 0.f;
  // The warp's 32 half rows are contiguous in elem_k (4 consecutive hexahedra x 8 rows x 192 bytes).  Written straight from
  // registers every STG.128 would scatter 32 16-byte pieces over 32 sectors (measured: the stores cost as many L1
  // wavefronts as all shared-memory traffic of the kernel), so the rows are transposed through the warp's own record area
  // (dead now; read only by this warp) and stored as 12 fully coalesced 512-byte writes.
  constexpr int WROW = EK_ROW + 4;                                   // padded row: 16-byte stores of 8 lanes hit distinct banks
  const int lane = threadIdx.x & 31;
  float* wst = s_rec + (size_t)(threadIdx.x >> 5) * 4 * HEX_STRIDE;  // records of this warp's 4 hexahedra
  __syncwarp();
  {
    float4* my = reinterpret_cast<float4*>(wst + lane * WROW);
#pragma unroll
    for (int v = 0; v < EK_ROW / 4; ++v) my[v] = make_float4(row[4 * v], row[4 * v + 1], row[4 * v + 2], row[4 * v + 3]);
  }
  __syncwarp();
  const int e_w = in.e_begin + blockIdx.x * EPB + (threadIdx.x >> 5) * 4;   // first hexahedron of the warp
  float4* gdst = reinterpret_cast<float4*>(elem_k + (size_t)e_w * 8 * EK_ROW);
#pragma unroll
  for (int i = 0; i < EK_ROW / 4; ++i) {
    const int g = i * 32 + lane;                                     // float4 index inside the warp's 6144 bytes
    const int rw = g / (EK_ROW / 4), c = g - rw * (EK_ROW / 4);
    if (e_w + (rw >> 3) < in.n_hex) gdst[g] = *reinterpret_cast<const float4*>(wst + rw * WROW + 4 * c);
  }
}

template <int kMode>
static size_t element_smem() {
  return (kMode != 0 ? (size_t)EPB * HEX_STRIDE : 4) * sizeof(float);
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
__global__ void __launch_bounds__(ETHREADS) elastic_energy_kernel(ElemIn in, int n_groups, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const int sub = threadIdx.x & 7;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int e = grp * EPB + (threadIdx.x >> 3);
    const bool live = e < in.n_hex;
    const int ee = live ? e : in.n_hex - 1;
    const int vtx = __ldg(in.hex + (size_t)ee * 8 + sub);
    float xn[3] = {__ldg(in.x + 3 * (size_t)vtx), __ldg(in.x + 3 * (size_t)vtx + 1), __ldg(in.x + 3 * (size_t)vtx + 2)};
    float X[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) X[a][c] = __shfl_sync(0xffffffffu, xn[c], a, 8);
    float fq[8][3];
    float en;
    gauss_point(in, ee, sub, X, nullptr, fq, en, false);
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

static ElemIn make_in(const FembLevel* l, const float* x, float mu, float la) {
  ElemIn in;
  in.hex = l->hex; in.x = x; in.invJ = l->invJ; in.detJ = l->detJ; in.e_begin = 0; in.n_hex = l->n_hex; in.mu = mu; in.la = la;
  return in;
}

static AsmArgs make_asm(const FembLevel* l, const float* diag_add, const float* elem_k, float* values) {
  AsmArgs A;
  A.n_vert = l->n_vert; A.i2v = l->i2v; A.v2i = l->v2i; A.vert_elem = l->vert_elem; A.elem2nnz = l->elem2nnz;
  A.L_ptr = l->L_ptr; A.U_ptr = l->U_ptr; A.off_l = l->off_l; A.off_d = l->off_d; A.off_u = l->off_u;
  A.fold = l->fold_plan; A.elem_k = elem_k; A.diag_add = diag_add; A.values = values;
  return A;
}

template <int kMode>
static int launch_element(ElemIn in, int e_begin, int e_end, float* elem_f, float* elem_k, cudaStream_t st) {
  static bool configured_dev[64] = {};                         // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  const size_t smem = element_smem<kMode>();
  if (!configured) {
    cudaFuncSetAttribute(element_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  if (e_end <= e_begin) return FEMB_OK;
  in.e_begin = e_begin; in.n_hex = e_end;
  element_kernel<kMode><<<cdiv(e_end - e_begin, EPB), ETHREADS, smem, st>>>(in, elem_f, elem_k);
  count_launch();
  return check_launch("element_kernel");
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

// element matrices of all hexahedra + fold of all rows; kMode 1 = Hessian only, 2 = element forces as well
template <int kMode>
static int hessian_pipeline(const FembLevel* l, const float* x, float mu, float la, const float* diag_add, float* elem_f,
                            float* elem_k, float* values, cudaStream_t st) {
  const ElemIn in = make_in(l, x, mu, la);
  const AsmArgs as = make_asm(l, diag_add, elem_k, values);
  // (A chunked two-stream software pipeline of the two stages -- the fold of chunk c next to the element kernel of chunk
  // c+1, element matrices read from L2 -- was built and measured: 1.88-2.10 vs 1.69-1.83 ms back to back, both stages are
  // bound on the SM side; removed)
  int rc = launch_element<kMode>(in, 0, l->n_hex, elem_f, elem_k, st);
  if (rc) return rc;
  return launch_assemble(as, 0, l->n_vert, st);
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_set_fold_config(int rows_per_warp, int min_blocks_per_sm) {
  if (rows_per_warp >= 1 && rows_per_warp <= 64) g_fold_rows_per_warp = rows_per_warp;
  if (min_blocks_per_sm == 4 || min_blocks_per_sm == 5) g_fold_min_blocks = min_blocks_per_sm;
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
    int rc = launch_element<0>(make_in(lvl, x, mu, lambda), 0, lvl->n_hex, elem_f, nullptr, as_stream(stream));
    if (rc) return rc;
  }
  if (lvl->n_vert == 0) return FEMB_OK;
  return launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, ordered, elem_f, out, as_stream(stream));
}

int femb_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* diag_add, float* elem_k,
                 float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && values, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  if (fused_available(lvl)) {
    FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ, "level not initialised");
    return launch_fused(lvl, x, mu, lambda, diag_add, values, nullptr, 0.f, nullptr, nullptr, 0.f, nullptr, as_stream(stream));
  }
  FEMB_REQUIRE(elem_k, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->elem2nnz && lvl->vert_elem && lvl->i2v && lvl->fold_plan, "level not initialised");
  return hessian_pipeline<1>(lvl, x, mu, lambda, diag_add, nullptr, elem_k, values, as_stream(stream));
}

int femb_force_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
                       const int* pin, const float* pin_pos, float control_mag, const float* diag_add, float* elem_f,
                       float* elem_k, float* force_out, float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && force_out && values, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  if (fused_available(lvl)) {
    FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ, "level not initialised");
    return launch_fused(lvl, x, mu, lambda, diag_add, values, mass, g, pin, pin_pos, control_mag, force_out, as_stream(stream));
  }
  FEMB_REQUIRE(elem_f && elem_k, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->elem2nnz && lvl->vert_elem && lvl->i2v && lvl->fold_plan && lvl->v2i,
               "level not initialised");
  int rc = hessian_pipeline<2>(lvl, x, mu, lambda, diag_add, elem_f, elem_k, values, as_stream(stream));
  if (rc) return rc;
  return launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, 1, elem_f, force_out, as_stream(stream));
}

int femb_energy(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g, const int* pin,
                const float* pin_pos, float control_mag, float* scratch, float* energy, void* stream) {
  FEMB_REQUIRE(lvl && x && scratch && energy, "null pointer");
  cudaStream_t st = as_stream(stream);
  const int n_groups = (int)cdiv(lvl->n_hex, EPB);
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
