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
// and parks it in shared memory; in phase B lane a accumulates the 8 blocks K[a][b] over the 8 Gauss
// points with the closed form
//     K_ab[j][i] = det * ( (g_a . S g_b) d_ji + mu (g_a . g_b)(F F^T)_ji + mu u_b[j] u_a[i] + la u_a[j] u_b[i] )
// (== d(P g_a det)_j / d x_b,i of quai_hexagon.py:86-106).  Element results go to a scratch array and
// are summed into the fixed [L;D;U] value layout by a gather over a precomputed inverse of the
// element->nnz map (segmented sum, fixed order, no atomics).
#include <cstdio>

#include "femb_common.cuh"

namespace femb {

constexpr int EPB = 16;            // hexahedra per CTA
constexpr int ETHREADS = EPB * 8;  // 128 threads

// per (hexahedron, Gauss point) record in shared memory, in floats; written and read as float4:
//   [0..5] det*mu * F F^T (symmetric: 00 01 02 11 12 22)   [6] det*mu   [7] det*la
//   [8 + 9 b ..] per node b: g_b[3], u_b[3], det * S g_b[3]                      (72 floats)
constexpr int QP_FFT = 0;
constexpr int QP_C = 6;
constexpr int QP_GUT = 8;
constexpr int QP_SIZE = 84;   // 21 float4: 16-byte accesses of 8 consecutive lanes hit distinct banks (21 is odd)
constexpr int FQ_SIZE = 25;   // per (hexahedron, Gauss point) nodal forces 8x3, padded

struct ElemIn {
  const int* hex;
  const float* x;
  const float* invJ;
  const float* detJ;
  int n_hex;
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
      float* o = R + QP_GUT + 9 * b;
      o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
#pragma unroll
      for (int r = 0; r < 3; ++r) o[3 + r] = F[3 * r] * g[0] + F[3 * r + 1] * g[1] + F[3 * r + 2] * g[2];
      o[6] = det * (S[0] * g[0] + S[1] * g[1] + S[2] * g[2]);
      o[7] = det * (S[1] * g[0] + S[3] * g[1] + S[4] * g[2]);
      o[8] = det * (S[2] * g[0] + S[4] * g[1] + S[5] * g[2]);
    }
  }
  if (want_rec) {
    float4* r4 = reinterpret_cast<float4*>(rec);
#pragma unroll
    for (int v = 0; v < 20; ++v) r4[v] = make_float4(R[4 * v], R[4 * v + 1], R[4 * v + 2], R[4 * v + 3]);
  }
}

// kMode: 0 = forces only, 1 = Hessian only, 2 = both
template <int kMode>
__global__ void __launch_bounds__(ETHREADS) element_kernel(ElemIn in, float* __restrict__ elem_f, float* __restrict__ elem_k) {
  constexpr bool kForce = (kMode != 1);
  constexpr bool kHess = (kMode != 0);
  extern __shared__ __align__(16) float smem[];
  float* s_rec = smem;                                                // [EPB][8][QP_SIZE]   (Hessian only)
  float* s_f = smem + (kHess ? EPB * 8 * QP_SIZE : 0);                // [EPB][8q][8a][3]    (force only)
  const int sub = threadIdx.x & 7;
  const int le = threadIdx.x >> 3;
  const int e = blockIdx.x * EPB + le;
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

  float fq[8][3];
  float energy;
  float* rec = s_rec + ((size_t)le * 8 + sub) * QP_SIZE;
  gauss_point(in, ee, sub, X, rec, fq, energy, kHess);
  if (kForce) {
    float* o = s_f + ((size_t)le * 8 + sub) * FQ_SIZE;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[a * 3 + c] = fq[a][c];
  }
  __syncthreads();
  if (kForce && live) {
    // lane a sums its node's force over the Gauss points in q order
    float f[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float* o = s_f + ((size_t)le * 8 + q) * FQ_SIZE + sub * 3;
      f[0] += o[0]; f[1] += o[1]; f[2] += o[2];
    }
    float* dst = elem_f + ((size_t)e * 8 + sub) * 3;
    dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2];
  }
  if (!kHess) return;

  // ---- phase B: lane a accumulates K[a][b], b = 0..7 --------------------------------------
  float K[8][9];
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int k = 0; k < 9; ++k) K[b][k] = 0.f;
  const bool has_la = in.la != 0.f;
#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const float* r = s_rec + ((size_t)le * 8 + q) * QP_SIZE;
    float G[80];                                           // the whole Gauss-point record, 20 broadcast LDS.128
    const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
    for (int v = 0; v < 20; ++v) {
      const float4 t4 = r4[v];
      G[4 * v] = t4.x; G[4 * v + 1] = t4.y; G[4 * v + 2] = t4.z; G[4 * v + 3] = t4.w;
    }
    const float* FFT = G + QP_FFT;
    const float cmu = G[QP_C], cla = G[QP_C + 1];
    const float* ga = r + QP_GUT + 9 * sub;                // own node: runtime index, scalar loads
    const float g_a[3] = {ga[0], ga[1], ga[2]};
    const float ua_mu[3] = {cmu * ga[3], cmu * ga[4], cmu * ga[5]};
    const float ua_la[3] = {cla * ga[3], cla * ga[4], cla * ga[5]};
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const float* gb = G + QP_GUT + 9 * b;
      const float g_b[3] = {gb[0], gb[1], gb[2]};
      const float u_b[3] = {gb[3], gb[4], gb[5]};
      const float s1 = g_a[0] * gb[6] + g_a[1] * gb[7] + g_a[2] * gb[8];       // det * g_a . S g_b
      const float s2 = g_a[0] * g_b[0] + g_a[1] * g_b[1] + g_a[2] * g_b[2];    // g_a . g_b
      // K[j][i] += d_ji s1 + (det mu FF^T)_ji s2 + det mu u_b[j] u_a[i] + det la u_a[j] u_b[i]
      K[b][0] += s1 + FFT[0] * s2 + u_b[0] * ua_mu[0];
      K[b][1] += FFT[1] * s2 + u_b[0] * ua_mu[1];
      K[b][2] += FFT[2] * s2 + u_b[0] * ua_mu[2];
      K[b][3] += FFT[1] * s2 + u_b[1] * ua_mu[0];
      K[b][4] += s1 + FFT[3] * s2 + u_b[1] * ua_mu[1];
      K[b][5] += FFT[4] * s2 + u_b[1] * ua_mu[2];
      K[b][6] += FFT[2] * s2 + u_b[2] * ua_mu[0];
      K[b][7] += FFT[4] * s2 + u_b[2] * ua_mu[1];
      K[b][8] += s1 + FFT[5] * s2 + u_b[2] * ua_mu[2];
      if (has_la) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int i = 0; i < 3; ++i) K[b][3 * j + i] += ua_la[j] * u_b[i];
      }
    }
  }
  if (live) {
    float4* dst = reinterpret_cast<float4*>(elem_k + ((size_t)e * 8 + sub) * 72);    // K[e][a][b][3][3], a = sub
#pragma unroll
    for (int v = 0; v < 18; ++v) {
      const int i0 = 4 * v;
      dst[v] = make_float4(K[i0 / 9][i0 % 9], K[(i0 + 1) / 9][(i0 + 1) % 9], K[(i0 + 2) / 9][(i0 + 2) % 9],
                           K[(i0 + 3) / 9][(i0 + 3) % 9]);
    }
  }
}

template <int kMode>
static size_t element_smem() {
  size_t f = 0;
  if (kMode != 1) f += (size_t)EPB * 8 * FQ_SIZE;
  if (kMode != 0) f += (size_t)EPB * 8 * QP_SIZE;
  return f * sizeof(float);
}

// Row-centric assembly, one warp per index-order row.  A row needs row `a` of the element matrices of its
// <= 8 incident hexahedra: 8 blocks = 288 contiguous, 16-byte aligned bytes each, fetched as 18 float4 by 18
// lanes (all 8 fetches in flight before the first is consumed).  The blocks are folded into the row's <= 27
// value slots through the element->nnz map in shared memory (fixed order: incidence order, then b; no
// atomics) and the row's L, D and U segments are written with coalesced stores.
constexpr int AS_WARPS = 8;
constexpr int AS_THREADS = AS_WARPS * 32;
constexpr int AS_SLOTS = 28;
__global__ void __launch_bounds__(AS_THREADS) assemble_rows_kernel(int n_vert, const int* __restrict__ i2v,
                                                                  const int* __restrict__ vert_elem,
                                                                  const int* __restrict__ elem2nnz,
                                                                  const int* __restrict__ L_ptr, const int* __restrict__ U_ptr,
                                                                  int off_l, int off_d, int off_u,
                                                                  const float* __restrict__ elem_k,
                                                                  const float* __restrict__ diag_add, float* __restrict__ values) {
  __shared__ float s_acc[AS_WARPS][AS_SLOTS * 9 + 4];
  __shared__ __align__(16) float s_stage[AS_WARPS][8][72];
  __shared__ int s_loc[AS_WARPS][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * AS_WARPS + w;
  if (r >= n_vert) return;                                   // warp-uniform
  const int v = __ldg(i2v + r);
  const int lp = __ldg(L_ptr + r), nl = __ldg(L_ptr + r + 1) - lp;
  const int up = __ldg(U_ptr + r), nu = __ldg(U_ptr + r + 1) - up;
  // incident hexahedra (e*8+a): lane j < 8 holds entry j
  const int my_ea = lane < 8 ? __ldg(vert_elem + (size_t)v * 8 + lane) : -1;
  float4 seg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ea = __shfl_sync(0xffffffffu, my_ea, j);
    seg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ea >= 0 && lane < 18) seg[j] = __ldg(reinterpret_cast<const float4*>(elem_k + (size_t)ea * 72) + lane);
  }
  // local slot of every (j,b): lanes handle 2 map entries each
  const bool fits = nl + 1 + nu <= AS_SLOTS;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = lane + 32 * h, j = t >> 3, b = t & 7;
    const int ea = __shfl_sync(0xffffffffu, my_ea, j);
    int loc = -1;
    if (ea >= 0) {
      const int s = __ldg(elem2nnz + (size_t)ea * 8 + b);
      loc = s < off_d ? (s - off_l - lp) : (s < off_u ? nl : nl + 1 + (s - off_u - up));
      if (!fits) loc = s;                                    // oversized rows keep the global slot (slow path below)
    }
    s_loc[w][t] = loc;
  }
  for (int i = lane; i < AS_SLOTS * 9; i += 32) s_acc[w][i] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (lane < 18) *reinterpret_cast<float4*>(&s_stage[w][j][4 * lane]) = seg[j];
  __syncwarp();
  if (fits) {
    for (int j = 0; j < 8; ++j) {
      if (s_loc[w][8 * j] < 0) continue;                     // hexahedron slot unused (boundary vertex); warp-uniform
#pragma unroll
      for (int h = 0; h < 3; ++h) {
        const int t = lane + 32 * h;
        if (t < 72) {
          const int b = t / 9, k = t - 9 * b;
          s_acc[w][s_loc[w][8 * j + b] * 9 + k] += s_stage[w][j][t];
        }
      }
      __syncwarp();
    }
    float* outL = values + ((size_t)off_l + lp) * 9;
    for (int i = lane; i < nl * 9; i += 32) outL[i] = s_acc[w][i];
    if (lane < 9) {
      float d = s_acc[w][nl * 9 + lane];
      if (diag_add != nullptr && (lane == 0 || lane == 4 || lane == 8)) d += __ldg(diag_add + r);
      values[((size_t)off_d + r) * 9 + lane] = d;
    }
    float* outU = values + ((size_t)off_u + up) * 9;
    for (int i = lane; i < nu * 9; i += 32) outU[i] = s_acc[w][(nl + 1) * 9 + i];
  } else {
    // rows with more than 27 blocks (not produced by hexahedral lattices): one slot at a time
    for (int part = 0; part < 3; ++part) {
      const int cnt = part == 0 ? nl : (part == 1 ? 1 : nu);
      for (int li = 0; li < cnt; ++li) {
        const int slot = part == 0 ? off_l + lp + li : (part == 1 ? off_d + r : off_u + up + li);
        if (lane < 9) {
          float sum = (part == 1 && diag_add != nullptr && (lane == 0 || lane == 4 || lane == 8)) ? __ldg(diag_add + r) : 0.f;
          for (int t = 0; t < 64; ++t)
            if (s_loc[w][t] == slot) sum += s_stage[w][t >> 3][(t & 7) * 9 + lane];
          values[(size_t)slot * 9 + lane] = sum;
        }
      }
    }
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
  in.hex = l->hex; in.x = x; in.invJ = l->invJ; in.detJ = l->detJ; in.n_hex = l->n_hex; in.mu = mu; in.la = la;
  return in;
}

template <int kMode>
static int launch_element(const FembLevel* l, const float* x, float mu, float la, float* elem_f, float* elem_k,
                          cudaStream_t st) {
  static bool configured = false;
  const size_t smem = element_smem<kMode>();
  if (!configured) {
    cudaFuncSetAttribute(element_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  element_kernel<kMode><<<cdiv(l->n_hex, EPB), ETHREADS, smem, st>>>(make_in(l, x, mu, la), elem_f, elem_k);
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

static int launch_assemble(const FembLevel* l, const float* diag_add, const float* elem_k, float* values, cudaStream_t st) {
  assemble_rows_kernel<<<cdiv(l->n_vert, AS_WARPS), AS_THREADS, 0, st>>>(l->n_vert, l->i2v, l->vert_elem, l->elem2nnz, l->L_ptr,
                                                                        l->U_ptr, l->off_l, l->off_d, l->off_u, elem_k, diag_add,
                                                                        values);
  count_launch();
  return check_launch("assemble_rows_kernel");
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_prepare_rest(int n_hex, const int* hex, const float* x_rest, float* detJ, float* invJ, void* stream) {
  FEMB_REQUIRE(n_hex >= 0 && hex && x_rest && detJ && invJ, "null pointer or negative size");
  if (n_hex == 0) return FEMB_OK;
  prepare_rest_kernel<<<cdiv((long long)n_hex * 8, 256), 256, 0, as_stream(stream)>>>(n_hex, hex, x_rest, detJ, invJ);
  count_launch();
  return check_launch("prepare_rest_kernel");
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
    int rc = launch_element<0>(lvl, x, mu, lambda, elem_f, nullptr, as_stream(stream));
    if (rc) return rc;
  }
  if (lvl->n_vert == 0) return FEMB_OK;
  return launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, ordered, elem_f, out, as_stream(stream));
}

int femb_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* diag_add, float* elem_k,
                 float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && elem_k && values, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->elem2nnz && lvl->vert_elem && lvl->i2v, "level not initialised");
  if (lvl->n_hex > 0) {
    int rc = launch_element<1>(lvl, x, mu, lambda, nullptr, elem_k, as_stream(stream));
    if (rc) return rc;
  }
  if (lvl->nnz == 0) return FEMB_OK;
  return launch_assemble(lvl, diag_add, elem_k, values, as_stream(stream));
}

int femb_force_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
                       const int* pin, const float* pin_pos, float control_mag, const float* diag_add, float* elem_f,
                       float* elem_k, float* force_out, float* values, void* stream) {
  FEMB_REQUIRE(lvl && x && elem_f && elem_k && force_out && values, "null pointer");
  FEMB_REQUIRE(lvl->hex && lvl->invJ && lvl->detJ && lvl->elem2nnz && lvl->vert_elem && lvl->i2v && lvl->vert_elem && lvl->v2i,
               "level not initialised");
  if (lvl->n_hex > 0) {
    int rc = launch_element<2>(lvl, x, mu, lambda, elem_f, elem_k, as_stream(stream));
    if (rc) return rc;
  }
  if (lvl->n_vert == 0) return FEMB_OK;
  int rc = launch_force_gather(lvl, x, mass, g, pin, pin_pos, control_mag, 1, elem_f, force_out, as_stream(stream));
  if (rc) return rc;
  return launch_assemble(lvl, diag_add, elem_k, values, as_stream(stream));
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
