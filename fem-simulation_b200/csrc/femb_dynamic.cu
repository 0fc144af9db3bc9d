// Implicit-Euler bookkeeping of the dynamic hexahedral simulation (one Newton iteration per frame).
//
// Reference: /root/reference/exp1/simulation/  hexagon.py:174-203 (compute_Inertia_Gradient_Kernel,
// Basic_Update_Kernel, Hessian_Mass_Kernel), cublas.py:234-299 (updateVelocity, Control_Kernel, Fixed_Update_Kernel,
// Hessian_Diag_Kernel, compute_partial_more_fixed_energy_X), object.py:1286-1325 (render).  The reference launches
// one tiny kernel per line; here each phase of the frame is one fused vertex kernel.
#include "femb_common.cuh"

namespace femb {

// Fixed_Update_Kernel + copy(old_x) + Basic_Update_Kernel + copy(inertia_x)          (object.py:1296-1299)
__global__ void dyn_begin_kernel(int n, float* __restrict__ x, float* __restrict__ v, float* __restrict__ old_x,
                                 float* __restrict__ inertia_x, float* __restrict__ fixed_x, const int* __restrict__ more_fixed,
                                 const int* __restrict__ pin, float d0, float d1, float d2, float damping, float dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dir[3] = {d0, d1, d2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t k = 3 * (size_t)i + c;
    const float xi = x[k];
    if (more_fixed[i] == 1 && pin[i] == 0) fixed_x[k] = xi + dir[c];
    old_x[k] = xi;
    const float vi = v[k] * damping;
    v[k] = vi;
    const float xn = xi + vi * dt;
    x[k] = xn;
    inertia_x[k] = xn;
  }
}

// Hessian_Diag_Kernel + Hessian_Mass_Kernel as the per-row diagonal add of the assembly     (object.py:1302-1304)
__global__ void dyn_diag_kernel(int n, const int* __restrict__ v2i, const float* __restrict__ mass,
                                const int* __restrict__ more_fixed, const int* __restrict__ pin, float control_mag, float inv_dt,
                                float* __restrict__ diag) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  float d = (more_fixed[v] == 1 || pin[v] == 1) ? control_mag : 0.f;
  d += mass[v] * inv_dt * inv_dt;
  diag[v2i[v]] = d;
}

// compute_partial_more_fixed_energy_X + compute_Inertia_Gradient_Kernel                      (object.py:1311-1312)
__global__ void dyn_force_kernel(int n, const int* __restrict__ v2i, const float* __restrict__ x,
                                 const float* __restrict__ inertia_x, const int* __restrict__ more_fixed,
                                 const float* __restrict__ fixed_x, const float* __restrict__ mass, float control_mag,
                                 float inv_dt, float* __restrict__ grad) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int r = v2i[v];
  const float c = mass[v] * inv_dt * inv_dt;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const size_t p = 3 * (size_t)v + k;
    float g = grad[3 * (size_t)r + k];
    if (more_fixed[v] == 1) g += control_mag * (fixed_x[p] - x[p]);
    g -= c * (x[p] - inertia_x[p]);
    grad[3 * (size_t)r + k] = g;
  }
}

// updateVelocity (cublas.py:234-237)
__global__ void dyn_velocity_kernel(int n3, const float* __restrict__ x, const float* __restrict__ old_x, float inv_dt,
                                    float* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) v[i] = (x[i] - old_x[i]) * inv_dt;
}

// Control_Kernel (cublas.py:275-280): vertices within sqrt(0.002) of the selected one become drag-constrained
__global__ void control_kernel(int n, int select_v, const float* __restrict__ x, int* __restrict__ more_fixed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d2 = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = x[3 * (size_t)i + c] - x[3 * (size_t)select_v + c];
    d2 += d * d;
  }
  more_fixed[i] = d2 < 0.002f ? 1 : 0;
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_dyn_begin(int n_vert, float* x, float* v, float* old_x, float* inertia_x, float* fixed_x, const int* more_fixed,
                   const int* pin, const float* dir_host, float damping, float dt, void* stream) {
  FEMB_REQUIRE(n_vert >= 0 && x && v && old_x && inertia_x && fixed_x && more_fixed && pin && dir_host, "null pointer");
  if (n_vert == 0) return FEMB_OK;
  dyn_begin_kernel<<<cdiv(n_vert, 256), 256, 0, as_stream(stream)>>>(n_vert, x, v, old_x, inertia_x, fixed_x, more_fixed, pin,
                                                                     dir_host[0], dir_host[1], dir_host[2], damping, dt);
  count_launch();
  return check_launch("dyn_begin_kernel");
}

int femb_dyn_diag(const FembLevel* lvl, const float* mass, const int* more_fixed, const int* pin, float control_mag,
                  float inv_dt, float* diag, void* stream) {
  FEMB_REQUIRE(lvl && mass && more_fixed && pin && diag, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  dyn_diag_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->v2i, mass, more_fixed, pin,
                                                                         control_mag, inv_dt, diag);
  count_launch();
  return check_launch("dyn_diag_kernel");
}

int femb_dyn_force(const FembLevel* lvl, const float* x, const float* inertia_x, const int* more_fixed, const float* fixed_x,
                   const float* mass, float control_mag, float inv_dt, float* grad, void* stream) {
  FEMB_REQUIRE(lvl && x && inertia_x && more_fixed && fixed_x && mass && grad, "null pointer");
  if (lvl->n_vert == 0) return FEMB_OK;
  dyn_force_kernel<<<cdiv(lvl->n_vert, 256), 256, 0, as_stream(stream)>>>(lvl->n_vert, lvl->v2i, x, inertia_x, more_fixed,
                                                                          fixed_x, mass, control_mag, inv_dt, grad);
  count_launch();
  return check_launch("dyn_force_kernel");
}

int femb_dyn_velocity(int n_vert, const float* x, const float* old_x, float inv_dt, float* v, void* stream) {
  FEMB_REQUIRE(n_vert >= 0 && x && old_x && v, "null pointer");
  if (n_vert == 0) return FEMB_OK;
  dyn_velocity_kernel<<<cdiv(3 * (long long)n_vert, 256), 256, 0, as_stream(stream)>>>(3 * n_vert, x, old_x, inv_dt, v);
  count_launch();
  return check_launch("dyn_velocity_kernel");
}

int femb_control(int n_vert, int select_v, const float* x, int* more_fixed, void* stream) {
  FEMB_REQUIRE(n_vert >= 0 && x && more_fixed && select_v >= 0 && select_v < n_vert, "bad argument");
  control_kernel<<<cdiv(n_vert, 256), 256, 0, as_stream(stream)>>>(n_vert, select_v, x, more_fixed);
  count_launch();
  return check_launch("control_kernel");
}

}  // extern "C"
