// Whole-step entry points of the C ABI: one call = one loop body of a solver method of the reference's `class Object`
// (object.py of exp1/quasi_simulation, exp1/simulation, exp1/cloth_simulation under /root/reference).  Each is the exact
// sequence of operator calls the Python drivers issue (fem-simulation_b200/object.py, object_dynamic.py, cloth.py), on one
// stream, with no allocation, no synchronisation and no read-back -- graph-capturable, and callable from any host language
// that can fill the context structs.
#include <nvtx3/nvToolsExt.h>

#include "femb_common.cuh"

namespace femb {

// NVTX ranges per stage of a step: host-side markers around the enqueue calls -- with ncu / Nsight attached the
// kernels launched inside a range carry its name (ncu --nvtx --nvtx-include "assembly/"); without a tool they cost nothing
// (nvtx3 is header-only and resolves to no-ops).
struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

#define FEMB_TRY(call)        \
  do {                        \
    const int rc_ = (call);   \
    if (rc_ != FEMB_OK) return rc_; \
  } while (0)

static int check_ctx(const FembContext* c, bool need_plans) {
  FEMB_REQUIRE(c && c->layers >= 1 && c->levels && c->values && c->B && c->B_fixed && c->delta_x && c->x_solved && c->R,
               "context arrays");
  FEMB_REQUIRE(c->x && c->grad && c->elem_f && c->elem_k, "context vectors");
  FEMB_REQUIRE(c->layers == 1 || (c->Ut_hat && c->Us_hat), "transfer operators");
  FEMB_REQUIRE(!need_plans || c->layers == 1 || (c->plans && c->t_values), "Galerkin plans");
  return FEMB_OK;
}

static inline size_t vec_bytes(const FembLevel& l) { return sizeof(float) * 3 * (size_t)l.n_vert; }

// PerformGaussSeidel(layer, 3) -> downSample -> VCycle(layer+1) -> upSample -> PerformGaussSeidel(layer, 3)   (object.py:1216-1226)
static int vcycle(const FembContext* c, int l, void* st) {
  const FembLevel* lv = &c->levels[l];
  static const char* const names[] = {"vcycle level 0", "vcycle level 1", "vcycle level 2", "vcycle level 3", "vcycle level 4+"};
  Range rg(names[l < 4 ? l : 4]);
  FEMB_TRY(femb_gauss_seidel(lv, c->values[l], c->B_fixed[l], c->delta_x[l], 3, c->not_spd, st));
  if (l < c->layers - 1) {
    const FembLevel* lc = &c->levels[l + 1];
    // downSample (object.py:1021-1027)
    FEMB_TRY(femb_copy(3 * lv->n_vert, c->delta_x[l], c->x_solved[l], st));
    FEMB_TRY(femb_residual(lv, c->values[l], c->B_fixed[l], c->delta_x[l], c->R[l], st));
    FEMB_TRY(femb_transfer_mv(&c->Ut_hat[l], 1.0f, c->R[l], 0.0f, c->B_fixed[l + 1], st));
    FEMB_TRY(vcycle(c, l + 1, st));
    // upSample (object.py:1029-1035)
    FEMB_TRY(femb_axpby(3 * lc->n_vert, 1.0f, c->delta_x[l + 1], 1.0f, c->x_solved[l + 1], st));
    FEMB_TRY(femb_transfer_mv(&c->Us_hat[l], 1.0f, c->x_solved[l + 1], 1.0f, c->x_solved[l], st));
    FEMB_TRY(femb_spmv(lv, c->values[l], FEMB_PART_ALL, -1.0f, c->x_solved[l], 1.0f, c->B_fixed[l], st));
  }
  if (l == c->layers - 1) return FEMB_OK;
  return femb_gauss_seidel(lv, c->values[l], c->B_fixed[l], c->delta_x[l], 3, c->not_spd, st);
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_vcycle(const FembContext* c, int layer, void* stream) {
  FEMB_TRY(check_ctx(c, false));
  FEMB_REQUIRE(layer >= 0 && layer < c->layers, "layer");
  return vcycle(c, layer, stream);
}

int femb_galerkin_chain(const FembContext* c, void* stream) {
  FEMB_TRY(check_ctx(c, true));
  Range rg("galerkin");
  for (int i = 1; i < c->layers; ++i)
    FEMB_TRY(femb_galerkin(&c->levels[i - 1], &c->levels[i], c->values[i - 1], &c->Ut_hat[i - 1], &c->Us_hat[i - 1], &c->plans[i - 1],
                           c->t_values[i - 1], 1, c->spd_value, c->values[i], stream));
  return FEMB_OK;
}

int femb_newton_mg_step(const FembContext* c, void* stream) {
  FEMB_TRY(check_ctx(c, true));
  const FembLevel* l0 = &c->levels[0];
  cudaStream_t st = as_stream(stream);
  Range step("newton_mg_step");
  {
    Range rg("assembly");
    FEMB_TRY(femb_force_hessian(l0, c->x, c->mu, c->la, c->mass0, c->g, c->pin, c->pin_pos, c->control_mag, c->diag_fixed, c->elem_f,
                                c->elem_k, c->grad, c->values[0], stream));                                 // object.py:1232-1236, 1249
    if (c->spd_level0) FEMB_TRY(femb_spd_blocks(c->values[0], l0->nnz, c->spd_value, stream));
  }
  FEMB_TRY(femb_galerkin_chain(c, stream));                                                                 // :1252-1264
  FEMB_TRY(femb_copy(3 * l0->n_vert, c->grad, c->B_fixed[0], stream));                                      // :1266
  FEMB_TRY(vcycle(c, 0, stream));                                                                           // :1290
  FEMB_TRY(femb_axpby(3 * l0->n_vert, 1.0f, c->delta_x[0], 1.0f, c->x_solved[0], stream));                  // finish(), :1037-1040
  for (int i = 1; i < c->layers; ++i) cudaMemsetAsync(c->x_solved[i], 0, vec_bytes(c->levels[i]), st);
  return femb_update_x(l0, c->x_solved[0], c->x, stream);                                                   // :1296
}

int femb_fas_step(const FembContext* c, int variant, float* x, void* stream) {
  FEMB_TRY(check_ctx(c, false));
  FEMB_REQUIRE(variant >= 1 && variant <= 3 && c->layers >= 2 && c->Ut_noOrder && (variant != 3 || c->Ut_noOrder_hat), "variant / operators");
  if (x == nullptr) x = c->x;
  const FembLevel* l0 = &c->levels[0];
  const FembLevel* l1 = &c->levels[1];
  cudaStream_t st = as_stream(stream);
  const bool spd = variant == 1, grav_pins = variant >= 2, fixed_diag = variant >= 2;
  const float* mass = grav_pins ? c->mass0 : nullptr;
  const int* pin = grav_pins ? c->pin : nullptr;
  const float* pin_pos = grav_pins ? c->pin_pos : nullptr;
  Range step("fas_step");
  // fine level (object.py:1468-1487)
  cudaMemsetAsync(c->B_fixed[1], 0, vec_bytes(*l1), st);
  {
    Range rg("assembly");
    FEMB_TRY(femb_force_hessian(l0, x, c->mu, c->la, mass, c->g, pin, pin_pos, c->control_mag, fixed_diag ? c->diag_fixed : nullptr,
                                c->elem_f, c->elem_k, c->B_fixed[0], c->values[0], stream));
    if (spd) FEMB_TRY(femb_spd_blocks(c->values[0], l0->nnz, c->spd_value, stream));
  }
  {
    Range rg("fine smoothing");
    FEMB_TRY(femb_gauss_seidel(l0, c->values[0], c->B_fixed[0], c->delta_x[0], 1, c->not_spd, stream));
    FEMB_TRY(femb_update_x(l0, c->delta_x[0], x, stream));
  }
  // coarse level (object.py:1489-1509)
  Range rg("coarse correction");
  FEMB_TRY(femb_transfer_mv(&c->Ut_noOrder[0], 1.0f, x, 0.0f, c->x_solved[1], stream));
  FEMB_TRY(femb_force(l0, x, c->mu, c->la, mass, c->g, pin, pin_pos, c->control_mag, 1, c->elem_f, c->R[0], stream));
  FEMB_TRY(femb_transfer_mv(&c->Ut_hat[0], 1.0f, c->R[0], 0.0f, c->B_fixed[1], stream));
  FEMB_TRY(femb_hessian(l1, c->x_solved[1], c->mu, c->la, nullptr, c->elem_k, c->values[1], stream));
  if (variant == 3) FEMB_TRY(femb_coarse_pin_penalty(l1, &c->Ut_noOrder_hat[0], c->pin, c->control_mag, c->values[1], stream));
  if (spd) FEMB_TRY(femb_spd_blocks(c->values[1], l1->nnz, c->spd_value, stream));
  FEMB_TRY(femb_jacobi(l1, c->values[1], c->B_fixed[1], c->delta_x[1], c->B[1], 1, c->not_spd, stream));
  FEMB_TRY(femb_transfer_mv(&c->Us_hat[0], 1.0f, c->delta_x[1], 0.0f, c->delta_x[0], stream));
  return femb_update_x(l0, c->delta_x[0], x, stream);
}

int femb_dynamic_step(const FembContext* c, const FembDynamic* d, void* stream) {
  FEMB_TRY(check_ctx(c, false));
  FEMB_REQUIRE(d && d->x && d->v && d->old_x && d->inertia_x && d->more_fixed && d->diag_dyn && d->cg_p && d->cg_ap && d->cg_state && d->dir_host,
               "dynamic arrays");
  const FembLevel* l0 = &c->levels[0];
  const int n = l0->n_vert;
  // Fixed_Update_Kernel moves the targets of the dragged vertices: all_pin_pos is written by this step
  FEMB_TRY(femb_dyn_begin(n, d->x, d->v, d->old_x, d->inertia_x, const_cast<float*>(c->pin_pos), d->more_fixed, c->pin, d->dir_host, d->damping, d->dt, stream));
  FEMB_TRY(femb_dyn_diag(l0, c->mass0, d->more_fixed, c->pin, c->control_mag, d->inv_dt, d->diag_dyn, stream));
  FEMB_TRY(femb_force_hessian(l0, d->x, c->mu, c->la, c->mass0, c->g, c->pin, c->pin_pos, c->control_mag, d->diag_dyn, c->elem_f, c->elem_k,
                              c->grad, c->values[0], stream));                                              // exp1/simulation/object.py:1302-1310
  FEMB_TRY(femb_dyn_force(l0, d->x, d->inertia_x, d->more_fixed, c->pin_pos, c->mass0, c->control_mag, d->inv_dt, c->grad, stream));
  FEMB_TRY(femb_copy(3 * n, c->grad, c->B_fixed[0], stream));
  FEMB_TRY(femb_cg(l0, c->values[0], c->B_fixed[0], c->delta_x[0], c->B[0], d->cg_p, d->cg_ap, d->cg_state, 5, 1e-5f, 0, stream));   // :1321
  FEMB_TRY(femb_update_x(l0, c->delta_x[0], d->x, stream));
  return femb_dyn_velocity(n, d->x, d->old_x, d->inv_dt, d->v, stream);
}

int femb_cloth_step(const FembCloth* k, void* stream) {
  FEMB_REQUIRE(k && k->level && k->edge && k->rest && k->x && k->v && k->old_x && k->inertia_x && k->mass && k->values && k->grad,
               "cloth arrays");
  const int n = k->level->n_vert;
  const float zero_dir[3] = {0.f, 0.f, 0.f};
  FEMB_TRY(femb_dyn_begin(n, k->x, k->v, k->old_x, k->inertia_x, k->x, k->zeros_i, k->zeros_i, zero_dir, k->damping, k->dt, stream));
  FEMB_TRY(femb_cloth_assemble(n, k->n_edge, k->edge, k->rest, k->k, k->x, k->mass, k->g, k->inertia_x, k->inv_dt, k->inc_ptr, k->inc,
                               k->off_p0p1, k->off_p1p0, k->d_slot, k->edge_h, k->edge_g, k->values, k->grad, stream));
  FEMB_TRY(femb_copy(3 * n, k->grad, k->B_fixed, stream));
  FEMB_TRY(femb_cg(k->level, k->values, k->B_fixed, k->delta_x, k->B, k->cg_p, k->cg_ap, k->cg_state, 5, 1e-5f, 0, stream));
  FEMB_TRY(femb_cloth_update(n, k->delta_x, k->x, 0, 65, stream));
  return femb_dyn_velocity(n, k->x, k->old_x, k->inv_dt, k->v, stream);
}

}  // extern "C"
