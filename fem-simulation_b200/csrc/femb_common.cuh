// Shared helpers for libfemb200 (sm_100a).  Device math that mirrors the reference's @wp.func bodies
// lives here so that every kernel uses one definition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "femb200.h"

namespace femb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define FEMB_REQUIRE(cond, what)                           \
  do {                                                     \
    if (!(cond)) {                                         \
      femb::set_error("%s: invalid argument: %s", __func__, what); \
      return FEMB_E_INVALID;                               \
    }                                                      \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Shape-function gradient table  shapeFuncGrad[a][q][c]
// (/root/reference/exp1/quasi_simulation/cpu_function.py:4-9, tables object.py:701-722).
// Node a and Gauss point q are both ordered 4*b0+2*b1+b2 with sign +1 when the bit is set; the
// abscissa is the literal 0.57735.  (1 +- G) needs one rounding, the product of two such factors a
// second one, the final *0.125 is exact -> identical to the reference's float32 table.
// ---------------------------------------------------------------------------------------------
#define FEMB_G 0.57735f
__host__ __device__ __forceinline__ constexpr float sfg_factor(int a, int q, int axis) {
  // 1 + s_axis(a) * sigma_axis(q) * G
  return (((a >> (2 - axis)) & 1) == ((q >> (2 - axis)) & 1)) ? (1.0f + FEMB_G) : (1.0f - FEMB_G);
}
__host__ __device__ __forceinline__ constexpr float sfg(int a, int q, int c) {
  const float s = ((a >> (2 - c)) & 1) ? 1.0f : -1.0f;
  const int d0 = (c == 0) ? 1 : 0;        // the two other axes, in the reference's multiplication order
  const int d1 = (c == 2) ? 1 : 2;
  return (s * sfg_factor(a, q, d0)) * sfg_factor(a, q, d1) * 0.125f;
}

// ---------------------------------------------------------------------------------------------
// solve3x3 (cublas.py:59-91): <=3 CG iterations on a 3x3 block, same early exits.  Falling off the
// end of the reference function after the third iteration is taken as `return x`.
// A is row-major.  Returns 1 if a `not spd` event (dot < 0) was seen.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int solve3x3(const float* A, const float b[3], float x[3]) {
  float r[3] = {b[0], b[1], b[2]};
  float r_norm = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  if (r_norm < 1e-10f) return 0;
  float p[3] = {r[0], r[1], r[2]};
  int bad = 0;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    float Ap[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Ap[i] = A[3 * i] * p[0] + A[3 * i + 1] * p[1] + A[3 * i + 2] * p[2];
    const float dot = p[0] * Ap[0] + p[1] * Ap[1] + p[2] * Ap[2];
    if (dot < 0.f) bad = 1;
    if (dot < 1e-10f) return bad;
    const float alpha = r_norm / dot;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
    }
    const float old = r_norm;
    r_norm = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (r_norm < 1e-10f) return bad;
    const float beta = r_norm / old;
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = r[i] + beta * p[i];
  }
  return bad;
}

// ---------------------------------------------------------------------------------------------
// spd_matrix33f (cublas.py:115-125) with wp.eig3 restated: cyclic approximate-Givens Jacobi
// eigen-analysis of McAdams et al. 2011 on the LOWER triangle (4 sweeps), V from the accumulated
// quaternion, d = diag(V^T A V); negative d -> value; A <- V diag(d) V^T.   (parity unpinned)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_conj(int x, int y, int z, float& s11, float& s21, float& s22, float& s31,
                                            float& s32, float& s33, float* qV) {
  float ch = 2.f * (s11 - s22);
  float sh = s21;
  const bool bsel = 5.828427124f * sh * sh < ch * ch;
  const float w = 1.0f / sqrtf(ch * ch + sh * sh);
  ch = bsel ? w * ch : 0.923879532f;
  sh = bsel ? w * sh : 0.3826834323f;
  const float scale = ch * ch + sh * sh;
  const float a = (ch * ch - sh * sh) / scale;
  const float b = (2.f * sh * ch) / scale;
  const float t11 = s11, t21 = s21, t22 = s22, t31 = s31, t32 = s32, t33 = s33;
  s11 = a * (a * t11 + b * t21) + b * (a * t21 + b * t22);
  s21 = a * (-b * t11 + a * t21) + b * (-b * t21 + a * t22);
  s22 = -b * (-b * t11 + a * t21) + a * (-b * t21 + a * t22);
  s31 = a * t31 + b * t32;
  s32 = -b * t31 + a * t32;
  s33 = t33;
  float tmp[3] = {qV[0] * sh, qV[1] * sh, qV[2] * sh};
  sh *= qV[3];
  qV[0] *= ch; qV[1] *= ch; qV[2] *= ch; qV[3] *= ch;
  qV[z] += sh;
  qV[3] -= tmp[z];
  qV[x] += tmp[y];
  qV[y] -= tmp[x];
  const float n11 = s22, n21 = s32, n22 = s33, n31 = s21, n32 = s31, n33 = s11;
  s11 = n11; s21 = n21; s22 = n22; s31 = n31; s32 = n32; s33 = n33;
}

__device__ __forceinline__ void spd_block(float* A, float value) {
  float s11 = A[0], s21 = A[3], s22 = A[4], s31 = A[6], s32 = A[7], s33 = A[8];
  float qV[4] = {0.f, 0.f, 0.f, 1.f};
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    jacobi_conj(0, 1, 2, s11, s21, s22, s31, s32, s33, qV);
    jacobi_conj(1, 2, 0, s11, s21, s22, s31, s32, s33, qV);
    jacobi_conj(2, 0, 1, s11, s21, s22, s31, s32, s33, qV);
  }
  const float x = qV[0], y = qV[1], z = qV[2], w = qV[3];
  float V[9];
  V[0] = 1.f - 2.f * (y * y + z * z); V[1] = 2.f * (x * y - w * z); V[2] = 2.f * (x * z + w * y);
  V[3] = 2.f * (x * y + w * z); V[4] = 1.f - 2.f * (x * x + z * z); V[5] = 2.f * (y * z - w * x);
  V[6] = 2.f * (x * z - w * y); V[7] = 2.f * (y * z + w * x); V[8] = 1.f - 2.f * (x * x + y * y);
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {   // d_k = v_k^T A v_k
    float t0 = A[0] * V[k] + A[1] * V[3 + k] + A[2] * V[6 + k];
    float t1 = A[3] * V[k] + A[4] * V[3 + k] + A[5] * V[6 + k];
    float t2 = A[6] * V[k] + A[7] * V[3 + k] + A[8] * V[6 + k];
    d[k] = V[k] * t0 + V[3 + k] * t1 + V[6 + k] * t2;
    if (d[k] < 0.f) d[k] = value;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = V[3 * i] * d[0] * V[3 * j] + V[3 * i + 1] * d[1] * V[3 * j + 1] + V[3 * i + 2] * d[2] * V[3 * j + 2];
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMA tensor store (cp.async.bulk.tensor through a CUtensorMap kernel parameter) ---------------------------------
// a 3-D box shared -> global as a bulk group (the issuing thread must wait for the group's shared-memory reads before the
// CTA exits).  (Tensor LOADS need the box to start on a 16-byte boundary of global memory -- a first coordinate that is
// not a multiple of 4 floats raises an illegal-instruction fault on B200, scripts/tma_probe.cu -- which is why the folds,
// whose windows start at arbitrary hexahedra, do not use them.)
__device__ __forceinline__ void tma_store_3d_saddr(const void* tmap, uint32_t src_saddr, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(src_saddr), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory become visible to the async proxy (TMA) after this fence + a barrier
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// block-wide sum; result valid in thread 0.  `red` must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

}  // namespace femb
