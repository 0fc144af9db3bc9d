// Stand-alone probe of the TMA tensor copies around the plane-major element-matrix scratch (a [360][S] array of floats here):
//   tma_probe 1 [c0 [promo [swizzle]]]  one 2-D load, box {32, 45}, issued by one thread; c0 = first coordinate (hexahedron),
//                                        promo = L2 promotion enum (0 none, 2 = 128 B), swizzle = enum (0 none, 3 = 128 B)
//   tma_probe 2                          8 lanes issue 4 loads each (boxes {36, 45} and {36, 9} from 16-byte aligned windows)
//                                        against their own mbarriers: the pattern of the TMA-staged fold that was not kept
//   tma_probe 3                          three 2-D stores, box {32, 120}, 128-byte swizzle, from a tile written with the element
//                                        kernel's address formula (the product uses the same tile through a 3-D view)
// Each mode verifies what arrived.  Measured on B200 (driver 580, CUDA 12.9): mode 1 with c0 = 36 or 64 passes, with c0 = 37
// the kernel dies with "an illegal instruction was encountered" whatever the swizzle / promotion -- a tensor load must start
// on a 16-byte boundary of global memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe scripts/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_saddr(const void* tmap, uint32_t src_saddr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src_saddr), "r"(c0), "r"(c1)
               : "memory");
}

// mode 1
__global__ void load_one(const __grid_constant__ CUtensorMap map, int c0, int c1, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* stage = reinterpret_cast<float*>(sm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 45 * 128);
  if (threadIdx.x == 0 && (smem_u32(sm) & 1023u)) printf("dynamic shared memory base %u not 1024-aligned\n", smem_u32(sm));
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 45 * 128);
    tma_load_2d(stage, &map, c0, c1, bar);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 45 * 32; i += blockDim.x) out[i] = stage[i];
}
// mode 2: windows of W = 36 hexahedra starting on 16-byte boundaries, every box on a 128-byte boundary of shared memory
constexpr int W = 36;
constexpr int BOX45_BYTES = (45 * W * 4 + 127) / 128 * 128, BOX9_BYTES = (9 * W * 4 + 127) / 128 * 128;
constexpr int SLOT_BYTES = BOX45_BYTES + 3 * BOX9_BYTES;
__global__ void load_eight(const __grid_constant__ CUtensorMap map45, const __grid_constant__ CUtensorMap map9, int base, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  float* stage = reinterpret_cast<float*>(sm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8 * SLOT_BYTES);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    for (int a = 0; a < 8; ++a) mbar_init(bar + a, 1);
  __syncthreads();
  if (w == 0 && lane < 8 && lane != 3) {                        // slot 3 stays empty (divergent issue, like a missing slot)
    mbar_expect_tx(bar + lane, 72 * W * 4);
    unsigned char* dst = sm + lane * SLOT_BYTES;
    tma_load_2d(dst, &map45, base + 4 * lane, lane * 45, bar + lane);
    for (int m = 1; m <= 3; ++m) tma_load_2d(dst + BOX45_BYTES + (m - 1) * BOX9_BYTES, &map9, base + 4 * lane, (((lane - m) & 7) * 5 + m) * 9, bar + lane);
  }
  for (int a = 0; a < 8; ++a) {
    if (a == 3) continue;
    mbar_wait(bar + a, 0);
  }
  for (int i = threadIdx.x; i < 8 * SLOT_BYTES / 4; i += blockDim.x) out[i] = (i / (SLOT_BYTES / 4) == 3) ? -1.f : stage[i];
}
// mode 3: thread (lp, sub) writes value(plane, hex) for its 45 planes and 2 hexahedra with the element kernel's formula
__global__ void store_tile(const __grid_constant__ CUtensorMap map, int e_cta) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int sub = threadIdx.x & 7, lp = threadIdx.x >> 3;
  const unsigned stg = (smem_u32(sm) + 3328u + 1023u) & ~1023u;
  const unsigned rbase = stg + (unsigned)sub * (45u * 128u) + (unsigned)(lp & 1) * 8u;
  const unsigned r7 = (unsigned)(sub * 45) & 7u, ch = (unsigned)(lp >> 1);
  for (int ki = 0; ki < 45; ++ki) {
    const unsigned addr = rbase + (unsigned)(ki * 128) + ((ch ^ ((r7 + (unsigned)ki) & 7u)) << 4);
    const int plane = sub * 45 + ki;
    const float v0 = (float)(plane * 1000 + e_cta + 2 * lp), v1 = v0 + 1.f;
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v0), "f"(v1) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 0; j < 3; ++j) tma_store_2d_saddr(&map, stg + (unsigned)(j * 120 * 128), e_cta, j * 120);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const size_t S = 512;   // hexahedra per plane (multiple of 32)
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !p) { printf("no encoder\n"); return 2; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  float* em;
  CK(cudaMalloc(&em, 360 * S * 4));
  std::vector<float> h(360 * S);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i / S) * 1000 + (i % S));
  CK(cudaMemcpy(em, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const cuuint32_t one[2] = {1, 1};
  const cuuint64_t d2[2] = {S, 360};
  const cuuint64_t s2[1] = {S * 4};
  const cuuint32_t b45[2] = {32, 45}, b9[2] = {(cuuint32_t)W, 9}, bst[2] = {32, 120}, b45w[2] = {(cuuint32_t)W, 45};
  CUtensorMap m45, m9, mst;
  CUresult r1 = enc(&m45, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, em, d2, s2, b45, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&m9, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, em, d2, s2, b9, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r3 = enc(&mst, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, em, d2, s2, bst, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode results %d %d %d\n", (int)r1, (int)r2, (int)r3);
  if (r1 || r2 || r3) return 2;
  float* out;
  CK(cudaMalloc(&out, 8 * SLOT_BYTES));
  int bad = 0;
  if (mode == 1) {
    // extra arguments: c0 (first hexahedron of the box), L2 promotion (0 none, 2 = 128 B), swizzle (0 none, 3 = 128 B)
    const int c0 = argc > 2 ? atoi(argv[2]) : 37, c1 = 90;
    const int promo = argc > 3 ? atoi(argv[3]) : 2, swz = argc > 4 ? atoi(argv[4]) : 0;
    CUresult r4 = enc(&m45, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, em, d2, s2, b45, one, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz,
                      (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("c0 %d promo %d swizzle %d encode %d\n", c0, promo, swz, (int)r4);
    load_one<<<1, 128, 45 * 128 + 64 + 1024>>>(m45, c0, c1, out);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(45 * 32);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    for (int r = 0; r < 45; ++r)
      for (int x = 0; x < 32; ++x) {
        const int chunk = swz == 3 ? ((x >> 2) ^ (r & 7)) : (x >> 2);        // 128-byte swizzle: 16-byte chunk ^= row & 7
        bad += o[r * 32 + chunk * 4 + (x & 3)] != (float)((c1 + r) * 1000 + c0 + x);
      }
  } else if (mode == 2) {
    const int base = 100;
    CUtensorMap m45w;
    CUresult r5 = enc(&m45w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, em, d2, s2, b45w, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r5);
    CK(cudaFuncSetAttribute(load_eight, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SLOT_BYTES + 64));
    load_eight<<<1, 288, 8 * SLOT_BYTES + 64>>>(m45w, m9, base, out);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(8 * SLOT_BYTES / 4);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    for (int a = 0; a < 8; ++a) {
      if (a == 3) continue;
      for (int r = 0; r < 72; ++r) {
        int plane, off;
        if (r < 45) { plane = a * 45 + r; off = r * W; }
        else { const int m = (r - 45) / 9 + 1; plane = (((a - m) & 7) * 5 + m) * 9 + (r - 45) % 9; off = BOX45_BYTES / 4 + (m - 1) * (BOX9_BYTES / 4) + ((r - 45) % 9) * W; }
        for (int x = 0; x < W; ++x) bad += o[a * (SLOT_BYTES / 4) + off + x] != (float)(plane * 1000 + base + 4 * a + x);
      }
    }
  } else {
    const int e_cta = 64;
    CK(cudaMemset(em, 0, 360 * S * 4));
    CK(cudaFuncSetAttribute(store_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 67328));
    store_tile<<<1, 128, 67328>>>(mst, e_cta);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h.data(), em, h.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t pl = 0; pl < 360; ++pl)
      for (size_t x = 0; x < S; ++x) {
        const float want = (x >= (size_t)e_cta && x < (size_t)e_cta + 32) ? (float)(pl * 1000 + x) : 0.f;
        bad += h[pl * S + x] != want;
      }
  }
  printf("mode %d: %s (%d mismatches)\n", mode, bad ? "FAIL" : "ok", bad);
  return bad != 0;
}
