// One-kernel Hessian (+ force) assembly of a level whose mesh is a subset of a regular lattice.
//
// Reference (paths relative to /root/reference/exp1/quasi_simulation/):
//   compute_elastic_hessian             quai_hexagon.py:69-106   (N_h*64 threads, 4608 float atomics / hexahedron)
//   compute_partial_elastic_energy_X    quai_hexagon.py:109-133  (+ gravity / pin springs :160-186)
//   MF_value = MF_value_fixed           object.py:1233
//
// Why a second assembly path: the two-kernel path of femb_element.cu writes the half-stored element matrices to HBM
// (1536 B per hexahedron) and reads them back in the row fold -- 2.9x the compulsory traffic and two SM-bound kernels.
// Here the element matrices never leave the SM:
//   * a CTA owns a column of 7x7 vertex columns of the lattice and marches through the vertex planes of a z-chunk; per
//     layer of lattice cells it evaluates the 8x8 hexahedra that touch its vertices (one halo cell around the tile is
//     recomputed: 64 hexahedra per 49 rows), 32 at a time;
//   * phase A (lane = Gauss point) and phase B (lane = node) are those of the element kernel, but the Gauss-point
//     records are stored with the points 2p and 2p+1 interleaved, so that every operand of phase B is a natural
//     64-bit register pair (value at 2p, value at 2p+1) and the whole of phase B runs on packed fp32 FMAs
//     (fma.rn.f32x2 -> SASS FFMA2): the two halves of an accumulator are the partial sums over the even and the odd
//     Gauss points;
//   * every lane adds its 5 blocks (and the 3 mirrored ones, transposed) straight into the shared-memory rows of the
//     two open vertex planes; hexahedra of one (x, y) parity class share no vertex, so two barrier-separated
//     sub-phases per round make the read-modify-writes race-free and the summation order fixed (no atomics);
//   * when a vertex plane has seen the cell layers below and above it, its rows leave in the reference's [L;D;U]
//     order through a per-row permutation (lattice offset -> sorted position), as coalesced stores, pin / mass
//     diagonal and the force (element forces + gravity + pin springs) included.
// Inputs of the next round (cell -> hexahedron -> vertices -> positions, rest data) are fetched while the current round
// computes, because one CTA per SM (197 KB of shared memory) leaves nobody else to hide that latency.
#include "femb_common.cuh"

namespace femb {
namespace fused {

constexpr int FT = FEMB_FUSED_TILE;            // vertex columns per tile edge
constexpr int FC = FT + 1;                     // lattice cells per tile edge (one halo cell on the low side)
constexpr int NROW = FT * FT;                  // rows of one vertex plane of a tile
constexpr int BLK = 12;                        // floats per 3x3 block in shared memory (9 + 3 pad: float4 accesses)
constexpr int ROWF = 27 * BLK;                 // floats per row: one block per lattice offset 9(dx+1)+3(dy+1)+(dz+1)
constexpr int REC_QP = 124;                    // floats of the record of a pair of Gauss points
constexpr int REC_HEX = 4 * REC_QP;            // floats of the records of a hexahedron
constexpr int RR_WORDS = FEMB_ROW_REC_BYTES / 4;
constexpr int THREADS = 256;
constexpr int EPR = THREADS / 8;               // hexahedra per round
static_assert(FC == 8 && EPR == 32, "the round / parity-class mapping assumes 8x8 cells and 32 hexahedra per round");
// record of Gauss points (2p, 2p+1), every field a float2 (even point, odd point):
//   [0..5] det mu F F^T (00 01 02 11 12 22)   [6] det mu   [7] det la   [8..13] det S (00 01 02 11 12 22)
//   [14 + 6 b ..] node b: g_b[3] = J^-T dN_b, u_b[3] = F g_b
constexpr int RF_FFT = 0, RF_CMU = 6, RF_CLA = 7, RF_S = 8, RF_NODE = 14;

constexpr size_t SMEM_FLOATS = (size_t)2 * NROW * ROWF + 2 * NROW * 4 + (size_t)EPR * REC_HEX + 2 * NROW * RR_WORDS + 2 * FC * FC;
constexpr size_t SMEM_BYTES = SMEM_FLOATS * 4;

struct Args {
  const int* hex; const float* x; const float* invJ; const float* detJ;
  float mu, la;
  int vx, vy, vz;                 // lattice vertex dims (vx, vy multiples of FT)
  const int* cell_elem;           // [(vz+1)][(vy+1)][(vx+1)]: hexahedron of lattice cell (i-1, j-1, k-1) or -1
  const unsigned char* row_rec;   // [vz][vy][vx][FEMB_ROW_REC_BYTES]
  int tiles_x, tiles_y, planes_per_chunk;
  const float* diag_add; float* values; int off_l, off_d, off_u;
  const float* mass; float g; const int* pin; const float* pin_pos; float control_mag; float* force_out;
};

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return d;
}
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// what phase A of one (hexahedron, Gauss point) needs from global memory, fetched one round ahead
struct Inputs {
  int vtx;          // vertex id of node `sub` of the hexahedron
  float xn[3];      // its position
  float iJ[9];      // rest data of Gauss point `sub`
  float det;
};
__device__ __forceinline__ void fetch_stage1(const Args& P, int e, int sub, Inputs& in) {
  const int ee = e < 0 ? 0 : e;                                  // dead lanes shadow hexahedron 0 (never consumed)
  in.vtx = __ldg(P.hex + (size_t)ee * 8 + sub);
  const float* ip = P.invJ + ((size_t)ee * 8 + sub) * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) in.iJ[k] = __ldg(ip + k);
  in.det = __ldg(P.detJ + (size_t)ee * 8 + sub);
}
__device__ __forceinline__ void fetch_stage2(const Args& P, Inputs& in) {
#pragma unroll
  for (int c = 0; c < 3; ++c) in.xn[c] = __ldg(P.x + 3 * (size_t)in.vtx + c);
}

// Phase A for Gauss point q: the record (interleaved with the partner point) and the nodal force contributions -P g_b det.
// Same arithmetic as gauss_point() of femb_element.cu (quai_hexagon.py:77-133).
__device__ __forceinline__ void phase_a(const float (&X)[8][3], const float (&iJ)[9], float det, float mu, float la, int q,
                                        float* __restrict__ rec, float (&fq)[8][3]) {
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
  float E[6];
  E[0] = 0.5f * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6] - 1.f);
  E[1] = 0.5f * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
  E[2] = 0.5f * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]);
  E[3] = 0.5f * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7] - 1.f);
  E[4] = 0.5f * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
  E[5] = 0.5f * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8] - 1.f);
  const float trE = E[0] + E[3] + E[5];
  float S[6];
  const float tm = 2.f * mu;
  S[0] = tm * E[0] + la * trE; S[1] = tm * E[1]; S[2] = tm * E[2];
  S[3] = tm * E[3] + la * trE; S[4] = tm * E[4]; S[5] = tm * E[5] + la * trE;
  float Pk[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Pk[3 * r + 0] = F[3 * r] * S[0] + F[3 * r + 1] * S[1] + F[3 * r + 2] * S[2];
    Pk[3 * r + 1] = F[3 * r] * S[1] + F[3 * r + 1] * S[3] + F[3 * r + 2] * S[4];
    Pk[3 * r + 2] = F[3 * r] * S[2] + F[3 * r + 1] * S[4] + F[3 * r + 2] * S[5];
  }
  const float cmu = det * mu;
  rec[2 * (RF_FFT + 0)] = cmu * (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
  rec[2 * (RF_FFT + 1)] = cmu * (F[0] * F[3] + F[1] * F[4] + F[2] * F[5]);
  rec[2 * (RF_FFT + 2)] = cmu * (F[0] * F[6] + F[1] * F[7] + F[2] * F[8]);
  rec[2 * (RF_FFT + 3)] = cmu * (F[3] * F[3] + F[4] * F[4] + F[5] * F[5]);
  rec[2 * (RF_FFT + 4)] = cmu * (F[3] * F[6] + F[4] * F[7] + F[5] * F[8]);
  rec[2 * (RF_FFT + 5)] = cmu * (F[6] * F[6] + F[7] * F[7] + F[8] * F[8]);
  rec[2 * RF_CMU] = cmu;
  rec[2 * RF_CLA] = det * la;
#pragma unroll
  for (int k = 0; k < 6; ++k) rec[2 * (RF_S + k)] = det * S[k];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = dN[b][0] * iJ[j] + dN[b][1] * iJ[3 + j] + dN[b][2] * iJ[6 + j];
#pragma unroll
    for (int r = 0; r < 3; ++r) fq[b][r] = -det * (Pk[3 * r] * g[0] + Pk[3 * r + 1] * g[1] + Pk[3 * r + 2] * g[2]);
    float* o = rec + 2 * (RF_NODE + 6 * b);
    o[0] = g[0]; o[2] = g[1]; o[4] = g[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) o[6 + 2 * r] = F[3 * r] * g[0] + F[3 * r + 1] * g[1] + F[3 * r + 2] * g[2];
  }
}

// Phase B for node a of one hexahedron: the 5 blocks K[a][(a+k) & 7], k = 0..4, summed over the 8 Gauss points.
//   K_ab[j][i] = sum_q  d_ji (t_a . g_b) + (det mu FF^T)_ji (g_a . g_b) + det mu u_b[j] u_a[i] + det la u_a[j] u_b[i],
//   t_a = det S g_a.  Every operand is a (even point, odd point) pair; the two halves are added at the end.
struct BlockAcc {
  float2 K[5][9];
  float2 ks[5];
};
__device__ __forceinline__ void phase_b_pair(const float* __restrict__ rq, int a, bool has_la, BlockAcc& A) {
  const float4* r4 = reinterpret_cast<const float4*>(rq);
  const float4 h0 = r4[0], h1 = r4[1], h2 = r4[2], h3 = r4[3], h4 = r4[4], h5 = r4[5], h6 = r4[6];
  const float2 FFT[6] = {lo2(h0), hi2(h0), lo2(h1), hi2(h1), lo2(h2), hi2(h2)};
  const float2 cmu = lo2(h3), cla = hi2(h3);
  const float2 Sd[6] = {lo2(h4), hi2(h4), lo2(h5), hi2(h5), lo2(h6), hi2(h6)};
  const float4* na = reinterpret_cast<const float4*>(rq + 2 * (RF_NODE + 6 * a));
  const float4 o0 = na[0], o1 = na[1], o2 = na[2];
  const float2 g_a[3] = {lo2(o0), hi2(o0), lo2(o1)};
  const float2 u_a[3] = {hi2(o1), lo2(o2), hi2(o2)};
  float2 t_a[3];
  t_a[0] = fma2(Sd[0], g_a[0], fma2(Sd[1], g_a[1], mul2(Sd[2], g_a[2])));
  t_a[1] = fma2(Sd[1], g_a[0], fma2(Sd[3], g_a[1], mul2(Sd[4], g_a[2])));
  t_a[2] = fma2(Sd[2], g_a[0], fma2(Sd[4], g_a[1], mul2(Sd[5], g_a[2])));
  const float2 ua_mu[3] = {mul2(cmu, u_a[0]), mul2(cmu, u_a[1]), mul2(cmu, u_a[2])};
  float2 ua_la[3];
  if (has_la) {
#pragma unroll
    for (int i = 0; i < 3; ++i) ua_la[i] = mul2(cla, u_a[i]);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float2 gb[3], ub[3];
    if (k == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { gb[i] = g_a[i]; ub[i] = u_a[i]; }
    } else {
      const float4* nb = reinterpret_cast<const float4*>(rq + 2 * (RF_NODE + 6 * ((a + k) & 7)));
      const float4 p0 = nb[0], p1 = nb[1], p2 = nb[2];
      gb[0] = lo2(p0); gb[1] = hi2(p0); gb[2] = lo2(p1);
      ub[0] = hi2(p1); ub[1] = lo2(p2); ub[2] = hi2(p2);
    }
    A.ks[k] = fma2(t_a[0], gb[0], fma2(t_a[1], gb[1], fma2(t_a[2], gb[2], A.ks[k])));
    const float2 s2 = fma2(g_a[0], gb[0], fma2(g_a[1], gb[1], mul2(g_a[2], gb[2])));
    A.K[k][0] = fma2(FFT[0], s2, fma2(ub[0], ua_mu[0], A.K[k][0]));
    A.K[k][1] = fma2(FFT[1], s2, fma2(ub[0], ua_mu[1], A.K[k][1]));
    A.K[k][2] = fma2(FFT[2], s2, fma2(ub[0], ua_mu[2], A.K[k][2]));
    A.K[k][3] = fma2(FFT[1], s2, fma2(ub[1], ua_mu[0], A.K[k][3]));
    A.K[k][4] = fma2(FFT[3], s2, fma2(ub[1], ua_mu[1], A.K[k][4]));
    A.K[k][5] = fma2(FFT[4], s2, fma2(ub[1], ua_mu[2], A.K[k][5]));
    A.K[k][6] = fma2(FFT[2], s2, fma2(ub[2], ua_mu[0], A.K[k][6]));
    A.K[k][7] = fma2(FFT[4], s2, fma2(ub[2], ua_mu[1], A.K[k][7]));
    A.K[k][8] = fma2(FFT[5], s2, fma2(ub[2], ua_mu[2], A.K[k][8]));
    if (has_la) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) A.K[k][3 * j + i] = fma2(ua_la[j], ub[i], A.K[k][3 * j + i]);
    }
  }
}

// shared-memory row (0 .. 2*NROW-1) of node `n` of tile cell (ci, cj), or -1 if that vertex is not a row of this CTA
__device__ __forceinline__ int row_of(int ci, int cj, int n, int slot_lo, bool lo_ok, bool hi_ok) {
  const int vi = ci - 1 + ((n >> 2) & 1), vj = cj - 1 + ((n >> 1) & 1), az = n & 1;
  if ((unsigned)vi >= (unsigned)FT || (unsigned)vj >= (unsigned)FT) return -1;
  if (az ? !hi_ok : !lo_ok) return -1;
  return (slot_lo ^ az) * NROW + vj * FT + vi;
}
// lattice offset slot of node b seen from node a: 9 (dx+1) + 3 (dy+1) + (dz+1)
__device__ __forceinline__ int offset_slot(int a, int b) {
  return 13 + 9 * (((b >> 2) & 1) - ((a >> 2) & 1)) + 3 * (((b >> 1) & 1) - ((a >> 1) & 1)) + ((b & 1) - (a & 1));
}
__device__ __forceinline__ void add_block(float* __restrict__ dst, const float (&k)[9], bool transposed) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  float4 v0 = d4[0], v1 = d4[1], v2 = d4[2];
  if (!transposed) {
    v0.x += k[0]; v0.y += k[1]; v0.z += k[2]; v0.w += k[3];
    v1.x += k[4]; v1.y += k[5]; v1.z += k[6]; v1.w += k[7];
  } else {
    v0.x += k[0]; v0.y += k[3]; v0.z += k[6]; v0.w += k[1];
    v1.x += k[4]; v1.y += k[7]; v1.z += k[2]; v1.w += k[5];
  }
  v2.x += k[8];
  d4[0] = v0; d4[1] = v1; d4[2] = v2;
}

template <bool kForce>
__global__ void __launch_bounds__(THREADS, 1) fused_assemble_kernel(const Args P) {
  extern __shared__ __align__(16) float smem[];
  float* rows = smem;                                             // [2][NROW][27][BLK]
  float* rowf = rows + (size_t)2 * NROW * ROWF;                   // [2][NROW][4] forces
  float* rec = rowf + 2 * NROW * 4;                               // [EPR][REC_HEX]
  unsigned* rrec = reinterpret_cast<unsigned*>(rec + (size_t)EPR * REC_HEX);   // [2][NROW][RR_WORDS]
  int* s_elem = reinterpret_cast<int*>(rrec + 2 * NROW * RR_WORDS);            // [2][FC*FC]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, sub = tid & 7, le = tid >> 3;
  const int n_tiles = P.tiles_x * P.tiles_y;
  const int tile = blockIdx.x % n_tiles, chunk = blockIdx.x / n_tiles;
  const int ti0 = (tile % P.tiles_x) * FT, tj0 = (tile / P.tiles_x) * FT;
  const int K0 = chunk * P.planes_per_chunk, K1 = min(K0 + P.planes_per_chunk, P.vz);
  if (K0 >= K1) return;

  for (int i = tid; i < (2 * NROW * ROWF + 2 * NROW * 4) / 4; i += THREADS) reinterpret_cast<float4*>(rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // the hexahedron of this lane in round r sits in tile cell (ci, cjb + r): the warps of even / odd index hold cells of
  // even / odd ci, the two rounds cells of even / odd cj -- the four (x, y) parity classes
  const int ci = 2 * ((tid >> 3) & 3) + (w & 1), cjb = 2 * (w >> 1);
  const int cxd = P.vx + 1, cyd = P.vy + 1;
  const bool has_la = P.la != 0.f;

  int kk = K0 - 1;
  if (tid < FC * FC)
    s_elem[((kk + 2) & 1) * (FC * FC) + tid] =
        __ldg(P.cell_elem + ((size_t)(kk + 1) * cyd + (tj0 + (tid >> 3))) * cxd + (ti0 + (tid & 7)));
  __syncthreads();
  int e_cur = s_elem[((kk + 2) & 1) * (FC * FC) + cjb * FC + ci];
  Inputs cur;
  fetch_stage1(P, e_cur, sub, cur);
  fetch_stage2(P, cur);

  for (; kk < K1; ++kk) {
    const int sl = (kk + 2) & 1, su = sl ^ 1;                    // shared-memory slots of vertex planes kk (lower) and kk+1 (upper)
    const bool lo_ok = kk >= K0, hi_ok = kk + 1 < K1;
    int e_nextlayer = -1;
    if (tid < FC * FC && hi_ok)
      e_nextlayer = __ldg(P.cell_elem + ((size_t)(kk + 2) * cyd + (tj0 + (tid >> 3))) * cxd + (ti0 + (tid & 7)));

#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      const int cj = cjb + r;
      const bool live = e_cur >= 0;
      // ---- phase A: lane `sub` = Gauss point ------------------------------------------------------------
      float f[3] = {0.f, 0.f, 0.f};
      {
        float X[8][3];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) X[a][c] = __shfl_sync(0xffffffffu, cur.xn[c], a, 8);
        float fq[8][3];
        if (live) {
          phase_a(X, cur.iJ, cur.det, P.mu, P.la, sub, rec + (size_t)le * REC_HEX + (sub >> 1) * REC_QP + (sub & 1), fq);
        } else {
#pragma unroll
          for (int b = 0; b < 8; ++b) fq[b][0] = fq[b][1] = fq[b][2] = 0.f;
        }
        if (kForce) {
          // reduce-scatter over the 8 lanes of the hexahedron (xor 4, 2, 1): lane a ends with the force on node a
          float h4[4][3], h2[2][3];
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
        }
      }
      __syncthreads();                                           // records complete; output of the previous plane finished
      if (r == 0 && hi_ok) {
        // row records of vertex plane kk+1 (it leaves at the end of the next layer): fire and forget
        for (int i = tid; i < NROW * RR_WORDS; i += THREADS) {
          const int rw = i / RR_WORDS, wd = i - rw * RR_WORDS;
          const int vj = rw / FT, vi = rw - vj * FT;
          const size_t lat = ((size_t)(kk + 1) * P.vy + (tj0 + vj)) * P.vx + (ti0 + vi);
          cp_async4(rrec + (size_t)(su * NROW + rw) * RR_WORDS + wd, P.row_rec + lat * FEMB_ROW_REC_BYTES + 4 * wd);
        }
      }
      // ---- inputs of the next round, part 1 -----------------------------------------------------------
      int e_nxt = -1;
      if (r == 0) e_nxt = s_elem[sl * (FC * FC) + (cjb + 1) * FC + ci];
      else if (hi_ok) e_nxt = s_elem[su * (FC * FC) + cjb * FC + ci];
      const bool have_next = (r == 0) || hi_ok;
      Inputs nxt = cur;
      if (have_next) fetch_stage1(P, e_nxt, sub, nxt);

      // ---- phase B: lane `sub` = node ----------------------------------------------------------------
      float Kf[5][9];
      if (live) {
        BlockAcc A;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          A.ks[k] = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 9; ++i) A.K[k][i] = make_float2(0.f, 0.f);
        }
        const float* rh = rec + (size_t)le * REC_HEX;
#pragma unroll 1
        for (int p = 0; p < 2; ++p) phase_b_pair(rh + p * REC_QP, sub, has_la, A);
        if (have_next) fetch_stage2(P, nxt);
#pragma unroll 1
        for (int p = 2; p < 4; ++p) phase_b_pair(rh + p * REC_QP, sub, has_la, A);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float s = A.ks[k].x + A.ks[k].y;
#pragma unroll
          for (int i = 0; i < 9; ++i) Kf[k][i] = A.K[k][i].x + A.K[k][i].y;
          Kf[k][0] += s; Kf[k][4] += s; Kf[k][8] += s;
        }
      } else if (have_next) {
        fetch_stage2(P, nxt);
      }
      if (r == 0 && tid < FC * FC) s_elem[su * (FC * FC) + tid] = e_nextlayer;

      // ---- scatter into the rows of the two open planes: even-ci warps, barrier, odd-ci warps -------------
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {
        if ((w & 1) == ph && live) {
          const int ra = row_of(ci, cj, sub, sl, lo_ok, hi_ok);
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int b = (sub + k) & 7;
            const int n = offset_slot(sub, b);
            if (ra >= 0) add_block(rows + (size_t)ra * ROWF + n * BLK, Kf[k], false);
            if (k >= 1 && k <= 3) {
              const int rb = row_of(ci, cj, b, sl, lo_ok, hi_ok);
              if (rb >= 0) add_block(rows + (size_t)rb * ROWF + (26 - n) * BLK, Kf[k], true);
            }
          }
          if (kForce && ra >= 0) {
            float* fr = rowf + ra * 4;
            fr[0] += f[0]; fr[1] += f[1]; fr[2] += f[2];
          }
        }
        if (ph == 1 && r == 1) cp_async_wait_all();
        __syncthreads();
      }
      e_cur = e_nxt;
      cur = nxt;
    }

    // ---- vertex plane kk is complete: its rows leave in [L;D;U] order ------------------------------------------
    if (lo_ok) {
      // per-row scalars that live in global memory are fetched for all rows of the warp at once
      float dadd = 0.f;
      {
        const int idx = w + 8 * lane;
        if (P.diag_add != nullptr && lane < 7 && idx < NROW) {
          const int r = (int)rrec[(size_t)(sl * NROW + idx) * RR_WORDS + 8];
          if (r >= 0) dadd = __ldg(P.diag_add + r);
        }
      }
      float fout = 0.f;
      int f_row = -1;
      if (kForce) {
        const int j = lane >> 2, c = lane & 3, idx = w + 8 * j;
        if (j < 7 && c < 3 && idx < NROW) {
          const unsigned* rr = rrec + (size_t)(sl * NROW + idx) * RR_WORDS;
          const int v = (int)rr[7], r = (int)rr[8];
          if (r >= 0) {
            fout = rowf[(sl * NROW + idx) * 4 + c];
            if (P.mass != nullptr && c == 1) fout += __ldg(P.mass + v) * P.g;                         // quai_hexagon.py:176-180
            if (P.pin != nullptr) {
              const int pn = __ldg(P.pin + v);
              const float d = __ldg(P.pin_pos + 3 * (size_t)v + c) - __ldg(P.x + 3 * (size_t)v + c);
              if (pn == 1) fout += P.control_mag * d;                                               // quai_hexagon.py:160-165
            }
            f_row = 3 * r + c;
            rowf[(sl * NROW + idx) * 4 + c] = 0.f;
          }
        }
      }
      int it = 0;
      for (int idx = w; idx < NROW; idx += 8, ++it) {
        const unsigned* rr = rrec + (size_t)(sl * NROW + idx) * RR_WORDS;
        const int r = (int)rr[8];
        const float da = __shfl_sync(0xffffffffu, dadd, it);
        if (r < 0) continue;                                     // no vertex at this lattice point (warp-uniform)
        const int lp = (int)rr[9], up = (int)rr[10];
        const int nl = (int)(rr[11] & 255u), nu = (int)((rr[11] >> 8) & 255u);
        const unsigned char* perm = reinterpret_cast<const unsigned char*>(rr);     // sorted position -> lattice offset slot
        float* rowb = rows + (size_t)(sl * NROW + idx) * ROWF;
        float* outL = P.values + ((size_t)P.off_l + lp) * 9;
        for (int i = lane; i < nl * 9; i += 32) {
          const int pos = (i * 57) >> 9, fl = i - 9 * pos;
          outL[i] = rowb[perm[pos] * BLK + fl];
        }
        if (lane < 9) {
          float d = rowb[perm[nl] * BLK + lane];
          if (lane == 0 || lane == 4 || lane == 8) d += da;
          P.values[((size_t)P.off_d + r) * 9 + lane] = d;
        }
        float* outU = P.values + ((size_t)P.off_u + up) * 9;
        for (int i = lane; i < nu * 9; i += 32) {
          const int pos = (i * 57) >> 9, fl = i - 9 * pos;
          outU[i] = rowb[perm[nl + 1 + pos] * BLK + fl];
        }
        __syncwarp();
        for (int i = lane; i < ROWF / 4; i += 32) reinterpret_cast<float4*>(rowb)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (kForce && f_row >= 0) P.force_out[f_row] = fout;
    }
  }
}

// ---- setup: per lattice point the row record ----------------------------------------------------------------------
// bytes 0..26: sorted position within the row (L blocks, D, U blocks) -> lattice offset slot; 27: pad;
// int32 words 7..11: vertex id, index-order row (-1: no vertex here), L_ptr[row], U_ptr[row], nL | nU << 8
__global__ void __launch_bounds__(256) row_rec_init_kernel(size_t n_lat, unsigned* __restrict__ rec) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_lat * RR_WORDS) return;
  const int wd = (int)(t % RR_WORDS);
  rec[t] = wd < 7 ? 0xffffffffu : (wd <= 8 ? 0xffffffffu : 0u);
}
__global__ void __launch_bounds__(256) row_rec_header_kernel(int n_vert, const int* __restrict__ vert_lat, const int* __restrict__ v2i,
                                                            const int* __restrict__ L_ptr, const int* __restrict__ U_ptr,
                                                            unsigned* __restrict__ rec) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int r = __ldg(v2i + v);
  const int lp = __ldg(L_ptr + r), nl = __ldg(L_ptr + r + 1) - lp, up = __ldg(U_ptr + r), nu = __ldg(U_ptr + r + 1) - up;
  unsigned* o = rec + (size_t)__ldg(vert_lat + v) * RR_WORDS;
  o[7] = (unsigned)v; o[8] = (unsigned)r; o[9] = (unsigned)lp; o[10] = (unsigned)up;
  o[11] = (unsigned)((nl & 255) | ((nu & 255) << 8));
}
__global__ void __launch_bounds__(256) row_rec_perm_kernel(int n_hex, const int* __restrict__ hex, const int* __restrict__ vert_lat,
                                                          const int* __restrict__ v2i, const int* __restrict__ L_ptr,
                                                          const int* __restrict__ U_ptr, const int* __restrict__ elem2nnz,
                                                          int off_l, int off_d, int off_u, unsigned char* __restrict__ rec,
                                                          int* __restrict__ bad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_hex * 8) return;
  const int a = (int)(t & 7);
  const int v = __ldg(hex + t);
  const int r = __ldg(v2i + v);
  const int lp = __ldg(L_ptr + r), nl = __ldg(L_ptr + r + 1) - lp, up = __ldg(U_ptr + r), nu = __ldg(U_ptr + r + 1) - up;
  if (nl + 1 + nu > 27) { *bad = 1; return; }
  unsigned char* o = rec + (size_t)__ldg(vert_lat + v) * FEMB_ROW_REC_BYTES;
  for (int b = 0; b < 8; ++b) {
    const int s = __ldg(elem2nnz + t * 8 + b);
    const int loc = s < off_d ? (s - off_l - lp) : (s < off_u ? nl : nl + 1 + (s - off_u - up));
    if (loc < 0 || loc >= 27) { *bad = 1; continue; }
    o[loc] = (unsigned char)offset_slot(a, b);               // every writer of a byte writes the same value
  }
}

// z-chunks: enough CTAs to fill the machine in whole waves, each chunk paying one extra cell layer
static int pick_chunks(int n_tiles, int vz) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best = 1;
  double best_cost = 1e300;
  for (int c = 1; c <= vz; ++c) {
    const int ppc = (vz + c - 1) / c;
    if (ppc < 2 && c > 1) break;
    const int chunks = (vz + ppc - 1) / ppc;
    const long long ctas = (long long)n_tiles * chunks;
    const double waves = (double)((ctas + sms - 1) / sms);
    const double cost = waves * (ppc + 1.5);                   // layers per CTA (+ start-up: zeroing, first fetch)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = chunks; }
  }
  return best;
}

}  // namespace fused

bool fused_available(const FembLevel* l) { return l->lat_cell_elem != nullptr && l->lat_row_rec != nullptr && l->lat_vz > 0; }

// values (+ force_out if not NULL) of a lattice level in one kernel
int launch_fused(const FembLevel* l, const float* x, float mu, float la, const float* diag_add, float* values,
                 const float* mass, float g, const int* pin, const float* pin_pos, float control_mag, float* force_out,
                 cudaStream_t st) {
  using namespace fused;
  Args P;
  P.hex = l->hex; P.x = x; P.invJ = l->invJ; P.detJ = l->detJ; P.mu = mu; P.la = la;
  P.vx = l->lat_vx; P.vy = l->lat_vy; P.vz = l->lat_vz;
  P.cell_elem = l->lat_cell_elem; P.row_rec = l->lat_row_rec;
  P.tiles_x = l->lat_vx / FT; P.tiles_y = l->lat_vy / FT;
  if (l->lat_vx % FT != 0 || l->lat_vy % FT != 0 || P.tiles_x < 1 || P.tiles_y < 1) {
    set_error("fused assembly: lattice dims %d x %d are not multiples of %d", l->lat_vx, l->lat_vy, FT);
    return FEMB_E_INVALID;
  }
  const int n_tiles = P.tiles_x * P.tiles_y;
  int chunks = l->lat_chunks > 0 ? l->lat_chunks : pick_chunks(n_tiles, P.vz);
  if (chunks > P.vz) chunks = P.vz;
  P.planes_per_chunk = (P.vz + chunks - 1) / chunks;
  chunks = (P.vz + P.planes_per_chunk - 1) / P.planes_per_chunk;
  P.diag_add = diag_add; P.values = values; P.off_l = l->off_l; P.off_d = l->off_d; P.off_u = l->off_u;
  P.mass = mass; P.g = g; P.pin = pin_pos ? pin : nullptr; P.pin_pos = pin_pos; P.control_mag = control_mag; P.force_out = force_out;
  // the attribute is per device and cheap to set: no process-wide "configured" flag
  cudaFuncSetAttribute(fused_assemble_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  cudaFuncSetAttribute(fused_assemble_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  const unsigned grid = (unsigned)(n_tiles * chunks);
  if (force_out != nullptr) fused_assemble_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(P);
  else fused_assemble_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(P);
  count_launch();
  return check_launch("fused_assemble_kernel");
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_build_lattice_plan(const FembLevel* lvl, const int* vert_lat, unsigned char* row_rec, int* bad, void* stream) {
  FEMB_REQUIRE(lvl && vert_lat && row_rec && bad, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->v2i && lvl->elem2nnz && lvl->L_ptr && lvl->U_ptr, "level not initialised");
  FEMB_REQUIRE(lvl->lat_vx > 0 && lvl->lat_vy > 0 && lvl->lat_vz > 0, "lattice dims not set");
  cudaStream_t st = as_stream(stream);
  const size_t n_lat = (size_t)lvl->lat_vx * lvl->lat_vy * lvl->lat_vz;
  fused::row_rec_init_kernel<<<cdiv((long long)(n_lat * fused::RR_WORDS), 256), 256, 0, st>>>(n_lat, reinterpret_cast<unsigned*>(row_rec));
  count_launch();
  if (lvl->n_vert > 0) {
    fused::row_rec_header_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, st>>>(lvl->n_vert, vert_lat, lvl->v2i, lvl->L_ptr, lvl->U_ptr,
                                                                        reinterpret_cast<unsigned*>(row_rec));
    count_launch();
  }
  if (lvl->n_hex > 0) {
    fused::row_rec_perm_kernel<<<cdiv((long long)lvl->n_hex * 8, 256), 256, 0, st>>>(lvl->n_hex, lvl->hex, vert_lat, lvl->v2i, lvl->L_ptr,
                                                                                   lvl->U_ptr, lvl->elem2nnz, lvl->off_l, lvl->off_d,
                                                                                   lvl->off_u, row_rec, bad);
    count_launch();
  }
  return check_launch("femb_build_lattice_plan");
}

}  // extern "C"
