// Mass-spring cloth: spring energy gradient and 3x3 Hessian blocks per edge, implicit-Euler assembly.
//
// Reference: /root/reference/exp1/cloth_simulation/triangle.py:5-79 (compute_rest_length, compute_elastic_gradient,
// compute_elastic_hessian, compute_Inertia_Gradient, compute_gravity_Gradient, Hessian_Mass_Kernel) and
// object.py:657-677 there (render).  The reference scatters with atomics (4 blocks + 2 vec3 per edge); here an edge
// kernel writes the two off-diagonal blocks (one writer each) and parks H_e / g_e, and a vertex kernel gathers the
// diagonal block and the force over the incident edges in a fixed order.
#include "femb_common.cuh"

namespace femb {

__global__ void cloth_rest_kernel(int n_edge, const int* __restrict__ edge, const float* __restrict__ x, float* __restrict__ rest) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edge) return;
  const int p0 = edge[2 * e], p1 = edge[2 * e + 1];
  const float d0 = x[3 * p0] - x[3 * p1], d1 = x[3 * p0 + 1] - x[3 * p1 + 1], d2 = x[3 * p0 + 2] - x[3 * p1 + 2];
  rest[e] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
}

// H = (k - a) I + b d d^T, a = k L / |d|, b = a / |d|^2 (triangle.py:47-64); g = k (L/|d| - 1) d (triangle.py:26-36)
__global__ void cloth_edge_kernel(int n_edge, const int* __restrict__ edge, const float* __restrict__ rest, float k,
                                  const float* __restrict__ x, const int* __restrict__ off_p0p1, const int* __restrict__ off_p1p0,
                                  float* __restrict__ edge_h, float* __restrict__ edge_g, float* __restrict__ values) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edge) return;
  const int p0 = edge[2 * e], p1 = edge[2 * e + 1];
  const float d[3] = {x[3 * p0] - x[3 * p1], x[3 * p0 + 1] - x[3 * p1 + 1], x[3 * p0 + 2] - x[3 * p1 + 2]};
  const float len = rest[e];
  const float len_new = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const float a = k * len / len_new;
  const float b = a / len_new / len_new;
  const float gl = k * (len / len_new - 1.0f);
  float* v01 = values + (size_t)off_p0p1[e] * 9;
  float* v10 = values + (size_t)off_p1p0[e] * 9;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    edge_g[3 * (size_t)e + i] = gl * d[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float h = (i == j ? (k - a) : 0.f) + b * d[i] * d[j];
      edge_h[9 * (size_t)e + 3 * i + j] = h;
      v01[3 * i + j] = -h;
      v10[3 * i + j] = -h;
    }
  }
}

// diagonal block = m/dt^2 I + sum H_e ; force = sum +-g_e + m g e_y - m/dt^2 (x - x_inertia)     (object.py:665-672)
__global__ void cloth_vertex_kernel(int n_vert, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                    const float* __restrict__ edge_h, const float* __restrict__ edge_g,
                                    const float* __restrict__ mass, float g, const float* __restrict__ x,
                                    const float* __restrict__ inertia_x, float inv_dt, const int* __restrict__ d_slot,
                                    float* __restrict__ values, float* __restrict__ grad) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  float H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float f[3] = {0, 0, 0};
  const float c = mass[v] * inv_dt * inv_dt;
  H[0] = c; H[4] = c; H[8] = c;                                              // Hessian_Mass_Kernel
  for (int p = inc_ptr[v]; p < inc_ptr[v + 1]; ++p) {
    const int ee = inc[p];                                                   // edge*2 + end
    const int e = ee >> 1;
    const float s = (ee & 1) ? -1.f : 1.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) H[q] += edge_h[9 * (size_t)e + q];
#pragma unroll
    for (int q = 0; q < 3; ++q) f[q] += s * edge_g[3 * (size_t)e + q];
  }
  f[1] += mass[v] * g;                                                       // compute_gravity_Gradient
#pragma unroll
  for (int q = 0; q < 3; ++q) f[q] -= c * (x[3 * (size_t)v + q] - inertia_x[3 * (size_t)v + q]);   // compute_Inertia_Gradient
  float* dv = values + (size_t)d_slot[v] * 9;
#pragma unroll
  for (int q = 0; q < 9; ++q) dv[q] = H[q];
#pragma unroll
  for (int q = 0; q < 3; ++q) grad[3 * (size_t)v + q] = f[q];
}

// update_deltaX_kernel of the cloth (cublas.py:22-26 there): vertices `skip0` and `skip1` (0 and 65) stay put
__global__ void cloth_update_kernel(int n_vert, const float* __restrict__ dx, float* __restrict__ x, int skip0, int skip1) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert || v == skip0 || v == skip1) return;
#pragma unroll
  for (int q = 0; q < 3; ++q) x[3 * (size_t)v + q] = x[3 * (size_t)v + q] + dx[3 * (size_t)v + q];
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_cloth_rest_length(int n_edge, const int* edge, const float* x, float* rest, void* stream) {
  FEMB_REQUIRE(n_edge >= 0 && edge && x && rest, "null pointer");
  if (n_edge == 0) return FEMB_OK;
  cloth_rest_kernel<<<cdiv(n_edge, 256), 256, 0, as_stream(stream)>>>(n_edge, edge, x, rest);
  count_launch();
  return check_launch("cloth_rest_kernel");
}

int femb_cloth_assemble(int n_vert, int n_edge, const int* edge, const float* rest, float k, const float* x, const float* mass,
                        float g, const float* inertia_x, float inv_dt, const int* inc_ptr, const int* inc,
                        const int* off_p0p1, const int* off_p1p0, const int* d_slot, float* edge_h, float* edge_g,
                        float* values, float* grad, void* stream) {
  FEMB_REQUIRE(n_vert >= 0 && n_edge >= 0 && edge && rest && x && mass && inertia_x && inc_ptr && inc && off_p0p1 && off_p1p0 &&
                   d_slot && edge_h && edge_g && values && grad, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (n_edge > 0) {
    cloth_edge_kernel<<<cdiv(n_edge, 256), 256, 0, st>>>(n_edge, edge, rest, k, x, off_p0p1, off_p1p0, edge_h, edge_g, values);
    count_launch();
  }
  if (n_vert > 0) {
    cloth_vertex_kernel<<<cdiv(n_vert, 256), 256, 0, st>>>(n_vert, inc_ptr, inc, edge_h, edge_g, mass, g, x, inertia_x, inv_dt,
                                                           d_slot, values, grad);
    count_launch();
  }
  return check_launch("cloth_assemble");
}

int femb_cloth_update(int n_vert, const float* dx, float* x, int skip0, int skip1, void* stream) {
  FEMB_REQUIRE(n_vert >= 0 && dx && x, "null pointer");
  if (n_vert == 0) return FEMB_OK;
  cloth_update_kernel<<<cdiv(n_vert, 256), 256, 0, as_stream(stream)>>>(n_vert, dx, x, skip0, skip1);
  count_launch();
  return check_launch("cloth_update_kernel");
}

}  // extern "C"
