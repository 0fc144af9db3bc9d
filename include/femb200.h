/*
 * femb200.h -- C ABI of libfemb200.so: the B200 (sm_100a) implementation of the hot path of
 * caixiao-0725/fem-simulation (implicit / quasi-static hexahedral FEM step with multigrid).
 *
 * The reference has no FFI of its own for this path: every operator is a `wp.launch(kernel, dim,
 * inputs)` or a `bsr_*` call over caller-owned device arrays.  Each entry point
 * below replaces one such operator (or a fused group of them) and cites the reference lines it
 * stands for; paths are relative to /root/reference/exp1/quasi_simulation/ unless noted.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All `const float*`/`float*`/`int*` arguments
 *     are DEVICE pointers owned by the caller unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Nothing allocates,
 *     nothing synchronises, nothing reads back: every call only enqueues kernels, so a sequence of
 *     calls can be captured in a CUDA graph.  (The femb_setup_*_host entry points are host code on host pointers and
 *     return when done; femb_setup_hierarchy_host alone allocates -- host arrays it owns until femb_setup_hierarchy_free.)
 *   - vectors are float[n*3] (vec3 per vertex); block values are float[nnz*9], row-major 3x3, in the
 *     reference's [L;D;U] layout (object.py:499-505): L blocks sorted by (row,col) at [off_l, off_d),
 *     the diagonal block of row r at off_d + r, U blocks sorted by (row,col) from off_u.
 *   - solver vectors are in colour-sorted *index* order, geometry (x, hex, mass, pins) in *vertex*
 *     order, exactly like the reference.
 *   - return value: 0 on success, a negative FEMB_E_* code otherwise (femb_last_error() gives text).
 *   - thread-compatible, not thread-safe.  Process-wide state: the last-error string (thread-local), the launch counter
 *     (atomic), the NCCL function table (written once by the first femb_comm_* call) and the TUNING KNOBS femb_set_pdl /
 *     femb_set_gs_threads_per_row / femb_set_gs_rows_per_cta / femb_set_fold_config / femb_set_assembly_async -- plain ints read at launch time, meant to
 *     be set once before use (benchmark scripts, tests); they change launch shapes, and (threads per row) rounding order,
 *     never results beyond that.  Kernel attributes are set once per device.
 */
#ifndef FEMB200_H
#define FEMB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FEMB_OK 0
#define FEMB_E_INVALID (-1)   /* bad argument (null pointer, negative size, bad enum) */
#define FEMB_E_CUDA (-2)      /* a CUDA runtime call or launch failed                 */
#define FEMB_E_NODEVICE (-3)  /* no usable sm_100 device                              */

#define FEMB_ABI_VERSION 5

/* hexahedra per CTA of the element kernel (16 pairs: a thread carries two hexahedra in packed fp32 registers) */
#define FEMB_ASM_CHUNK_HEXES 32
#define FEMB_FOLD_SLOTS 27
#define FEMB_FOLD_PLAN_BYTES 128   /* 28*4 contributor bytes + 16 header bytes per vertex: one cache line */
/* floats of element-matrix scratch per hexahedron (elem_k of femb_hessian; the hexahedron count is rounded up to a multiple
 * of 32): the symmetric half of the 24x24 element matrix.  General layout: 8 rows x (5 blocks of 9 + 3 pad) per
 * hexahedron; lattice layout: blocks of 32 hexahedra, each 360 planes (row a, block k, component c) x 32 floats = 46 080
 * contiguous bytes (elem_k must be 16-byte aligned).  See femb_element.cu */
#define FEMB_ELEM_K_FLOATS 384
/* bytes of a row record of the lattice fold (FembLevel.lat_row_rec) */
#define FEMB_ROW_REC_BYTES 48

/* matrix part selector for femb_spmv (bsr_mv on self.L / self.D / self.U / self.A) */
#define FEMB_PART_L 1
#define FEMB_PART_D 2
#define FEMB_PART_U 4
#define FEMB_PART_ALL 7

/* One multigrid level: sizes + device pointers, filled once by the host after setup and read-only
 * afterwards.  Mirrors the per-level attributes built by Object.__init__ (object.py:116-865). */
typedef struct FembLevel {
  int n_vert;            /* dims[l]                                                     */
  int n_hex;             /* hexs[l].shape[0]                                            */
  int nnz;               /* MF_nnz / UtAUs_nnz[l-1]                                     */
  int off_l, off_d, off_u; /* object.py:503-505 / :617-619                              */
  int color_start[9];    /* color_vertex_num[l] (object.py:151-158, :234-241)           */
  const int* hex;        /* [n_hex*8] vertex ids, local node 4bx+2by+bz (object.py:34)  */
  const int* v2i;        /* [n_vert] dev_vertex2index[l]                                */
  const int* i2v;        /* [n_vert] dev_index2vertex[l]                                */
  const int* L_ptr;      /* [n_vert+1] row pointers into the L section                  */
  const int* L_col;      /* [L_nnz]  == MF_L_col / UtAUs_L_col                          */
  const int* U_ptr;      /* [n_vert+1] row pointers into the U section                  */
  const int* U_col;      /* [U_nnz]  == MF_U_col / UtAUs_U_col                          */
  const int* elem2nnz;   /* [n_hex*64] == hex_update_offset[l] (object.py:563-569)      */
  const int* vert_elem;  /* [n_vert*8] e*8+a of the hexes incident to a vertex, -1 pad  */
  const float* invJ;     /* [n_hex*8*9] inverse_pX_peps (quai_hexagon.py:18)            */
  const float* detJ;     /* [n_hex*8]   det_pX_peps     (quai_hexagon.py:17)            */
  const unsigned char* fold_plan; /* [n_vert*FEMB_FOLD_PLAN_BYTES] per vertex: for each of the 27 blocks of its row (L blocks, */
                                  /* D, U blocks) the <= 4 contributors j*8+p (j-th incident hexahedron, rotated column      */
                                  /* position p; 255 = none) in fold order, a 28th word with contributors 5..8 of the         */
                                  /* diagonal block, then int32 {row, L_ptr[row], U_ptr[row], nL | nU<<8}; filled by          */
                                  /* femb_build_fold_plan from vert_elem + elem2nnz                                          */
  /* Optional lattice fold (NULL = none): the mesh is a subset of a regular lattice with the reference's local node order
   * (every voxel mesh of the reference, object.py:28-45, and every coarse mesh built from one, object.py:171-241), so a row
   * has at most 27 blocks, one per lattice offset of the column vertex.  With it femb_hessian / femb_force_hessian keep
   * the element matrices plane-major and fold them with compile-time offsets (fold_lattice_kernel); without it they use
   * fold_plan.  Both filled by femb_build_lattice_fold. */
  const int* lat_inc8;            /* [n_vert*8] hexahedron in which the vertex is local node a, -1 = none               */
  const unsigned char* lat_row_rec; /* [n_vert*FEMB_ROW_REC_BYTES] per vertex: bytes 0..26 lattice offset                 */
                                  /* 9(dx+1)+3(dy+1)+dz+1 -> sorted position in the row (L blocks, D, U blocks; 255 = no */
                                  /* such block), int32 words 7..11 {vertex, row, L_ptr[row], U_ptr[row], nL | nU<<8}     */
} FembLevel;

/* CSR matrix with one scalar weight per entry = BSR matrix with blocks w*I3: the interpolation
 * operators Ut/Us/Ut_hat/Us_hat/Ut_noOrder/... (object.py:247-443). */
typedef struct FembTransfer {
  int nrow, ncol, nnz;
  const int* offsets;    /* [nrow+1] */
  const int* columns;    /* [nnz]    */
  const float* weights;  /* [nnz]    */
} FembTransfer;

/* Plan of the Galerkin product A_c = Ut_hat * (A_f * Us_hat) on fixed patterns
 * (replaces the per-step symbolic bsr_mm, sparse.py:950-1140, object.py:1259). */
typedef struct FembGalerkinPlan {
  int n_fine, n_coarse;
  int t_nnz;             /* blocks of T = A_f * Us_hat                                   */
  const int* t_ptr;      /* [n_fine+1]                                                   */
  const int* t_col;      /* [t_nnz] coarse index-order columns, sorted within a row      */
  const int* c_ptr;      /* [n_coarse+1] (row,col)-sorted pattern of A_c                 */
  const int* c_col;      /* [nnz_c]                                                      */
  const int* c_slot;     /* [nnz_c] [L;D;U] slot of each sorted block = inverse of       */
                         /*         UtAUs_block_offset (object.py:633-683)               */
  const unsigned char* t_slot;  /* [nnz_f*8] fine block s (in [L;D;U] slot order, column j), p-th parent */
                                /*           of j -> bits 0-4: position of that parent in T row(s),      */
                                /*           bits 5-6: e with weight 2^-e (used if dyadic); 255 = none   */
  const unsigned char* c_slot8; /* [t_nnz*8] T block t=(i,J), q-th parent I of i -> position of J in     */
                                /*           the sorted row I of A_c; 255 = none                         */
  const unsigned char* ut_q;    /* [nnz(Ut_hat)] entry (I,i) -> q with parent_q(i) == I                   */
  int dyadic;                   /* 1: every Us_hat weight is 1, 1/2, 1/4 or 1/8 and is encoded in t_slot  */
} FembGalerkinPlan;

/* Alignment contract of the streaming kernels: block-value arrays and the L_col / U_col arrays are
 * staged with 16-byte-granular bulk copies (cp.async.bulk), so their base must be 16-byte aligned and
 * they must be readable up to 16 bytes past their logical end (allocate 4 extra elements). */

/* ---- library ------------------------------------------------------------------------------ */
int femb_abi_version(void);
const char* femb_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long femb_launch_count(void);
/* 0 if device `device` exists and is compute capability 10.x */
int femb_check_device(int device);

/* ---- setup of the fixed patterns on the HOST (csrc/femb_setup.cu, OpenMP; every pointer is a host pointer) ----------------
 * Object.__init__ builds these maps with per-element Python loops and dictionaries (object.py:485-697) and bsr_mm repeats
 * the symbolic product every step (sparse.py:1018-1091); here they are built once, with the reference's numbering. */
/* vert_elem[n_vert*8]: e*8+a of the hexahedra incident to every vertex, ascending, -1 padded (the inverse of `hex`) */
int femb_setup_incidence_host(int n_vert, int n_hex, const int* hex, int* vert_elem);
/* CSR (offsets [nrow+1], columns / weights [nnz], sorted by (row, col)) of nnz unique COO entries: the scalar-weight form of
 * bsr_set_from_triplets (sparse.py:173-278) used for the interpolation operators and their transposes (object.py:247-443) */
int femb_setup_coo_to_csr_host(int nrow, int nnz, const int* rows, const int* cols, const float* w, int* offsets, int* columns,
                               float* weights);
/* the same with a second weight array on the same pattern (w2 / weights2 may both be NULL): the normalised and the
 * un-normalised interpolation weights of object.py:422-436 share rows and columns, one sort serves both */
int femb_setup_coo_to_csr_pair_host(int nrow, int nnz, const int* rows, const int* cols, const float* w, const float* w2, int* offsets,
                                    int* columns, float* weights, float* weights2);
/* pattern of one level = sorted set {(v2i[a], v2i[b]) : a, b in one hexahedron} in the [L;D;U] layout (object.py:485-599):
 * first the row pointers L_ptr / U_ptr [n_vert+1] (L_ptr[n_vert] = L_nnz = U_nnz), then, into arrays the caller sizes from
 * them, the columns, hex_update_offset (elem2nnz [n_hex*64], object.py:563-569) and UtAUs_block_offset (block_offset
 * [2 L_nnz + n_vert]: [L;D;U] slot -> rank in (row,col)-sorted order, object.py:633-683) */
int femb_setup_pattern_count_host(int n_vert, const int* hex, const int* v2i, const int* i2v, const int* vert_elem, int* L_ptr,
                                  int* U_ptr);
int femb_setup_pattern_fill_host(int n_vert, int n_hex, const int* hex, const int* v2i, const int* i2v, const int* vert_elem,
                                 const int* L_ptr, const int* U_ptr, int* L_col, int* U_col, int* elem2nnz, int* block_offset);
/* ---- the hierarchy itself (csrc/femb_setup_hierarchy.cu; host pointers): the dictionaries of Object.__init__ as ordered scans
 * over lattice lookup tables, numbering identical to the reference's first-seen order.  float arguments are float32 values of
 * the reference's spacings (dx, 2 dx, ...), `origin` [3] is the lattice origin (the reference: the bounding-box minimum). */
/* object.py:283-412: slot 9 i + 3 j + k of a coarse cell -> number of coarse corners cnt [27], their local numbers
 * corners [27*8] (ascending, zero padded) and the un-normalised weight [27] (1, 1/2, 1/4, 1/8) */
int femb_setup_stencil_host(int* cnt, int* corners, float* weight);
/* object.py:147-158: colour of vertex v = parity of its lattice index (i + 2 j + 4 k), index order = stable sort by colour:
 * v2i / i2v [n_vert] (vertex2index / index2vertex), color_vertex_num [9] */
int femb_setup_color_order_host(int n_vert, const float* x, const float* origin, float spacing, int* v2i, int* i2v,
                                int* color_vertex_num);
/* object.py:171-241, one coarsening step, in two calls.  count: hex_parent [n_hex] = coarse hexahedron of every hexahedron,
 * coarse_key (scratch the caller sizes [n_hex*3]; on return its first n_coarse_hex*3 entries are the lattice cells of the
 * coarse hexahedra in id order) and the two counts.  fill, into arrays sized from the counts: coarse_hex [n_coarse_hex*8],
 * coarse_x [n_coarse_vert*3], fine_coarse [n_coarse_hex*27] (-1 = no vertex), and the coarse level's v2i / i2v
 * [n_coarse_vert] and color_vertex_num [9] */
int femb_setup_coarsen_count_host(int n_vert, int n_hex, const float* x, const int* hex, const float* origin, float coarse_size,
                                  int* hex_parent, int* coarse_key, int* n_coarse_hex, int* n_coarse_vert);
int femb_setup_coarsen_fill_host(int n_vert, int n_hex, const float* x, const int* hex, const float* origin, float fine_size,
                                 float coarse_size, const int* hex_parent, const int* coarse_key, int n_coarse_hex,
                                 int n_coarse_vert, int* coarse_hex, float* coarse_x, int* fine_coarse, int* v2i, int* i2v,
                                 int* color_vertex_num);
/* object.py:265-420, the interpolation entries between two levels, in two calls.  count: first_pos [n_fine] = the first
 * (coarse hexahedron * 27 + slot) that names each fine vertex (-1: none) and the number of entries.  fill, entries in
 * ascending first_pos order: coarse_id / fine_id [n_entries] (vertex numbers: rowInd_noOrder / colInd_noOrder), entry_k
 * [n_entries] (position of the coarse vertex among the fine vertex's parents), weight (Ut_hat values), norm [n_coarse]
 * (column sums) and weight_norm = weight / norm[coarse_id] (Ut values) */
int femb_setup_interp_count_host(int n_fine, int n_coarse_hex, const int* fine_coarse, int* first_pos, int* n_entries);
int femb_setup_interp_fill_host(int n_fine, int n_coarse, int n_coarse_hex, const int* fine_coarse, const int* coarse_hex,
                                const int* first_pos, int n_entries, int* coarse_id, int* fine_id, unsigned char* entry_k,
                                float* weight, float* norm, float* weight_norm);
/* voxel set -> hexahedral mesh in the reference's numbering (object.py:28-45: lattice points x fastest restricted to the used
 * ones, cells x fastest, local node a = 4 bx + 2 by + bz).  mask [nx*ny*nz] bytes (z fastest, as a C-ordered [nx][ny][nz]
 * array), count first, then points [n_vert*3] = (lattice index + index_offset) * dx + origin (float32) and hex [n_hex*8] */
int femb_setup_mask_mesh_count_host(int nx, int ny, int nz, const unsigned char* mask, int* n_vert, int* n_hex);
int femb_setup_mask_mesh_fill_host(int nx, int ny, int nz, const unsigned char* mask, float dx, const float* origin,
                                   const int* index_offset, int n_vert, int n_hex, float* points, int* hex);
/* Wavefront OBJ surface meshes (object.py:28 calls pv.read): the "v x y z" and "f a b c ..." lines of `path` (indices 1-based or
 * negative = relative, "a/b/c" forms accepted, polygons split into fans).  count first, then read into verts [n_vert*3] (float64)
 * and faces [n_tri*3] (int64, 0-based) */
int femb_setup_obj_count_host(const char* path, long long* n_vert, long long* n_tri);
int femb_setup_obj_read_host(const char* path, long long n_vert, long long n_tri, double* verts, long long* faces);
/* the inside test of the voxeliser (object.py:28-34 calls pv.read + pv.voxelize; voxelize.py): inside [nx*ny*nz] (bytes, 0 / 1,
 * z fastest) = parity of the crossings of the +z ray from lattice point (gx[i], gy[j], gz[k]) with the n_face triangles
 * tri [n_face*9] (three xyz corners each, float64); the grids are strictly increasing; half-open edge rule */
int femb_setup_points_inside_host(int n_face, const double* tri, int nx, const double* gx, int ny, const double* gy, int nz,
                                  const double* gz, unsigned char* inside);
/* Object.__init__ (object.py:116-697) in one call: all levels of the hierarchy on the host, in arrays the library owns
 * (femb_setup_hierarchy_free releases them).  level[l] describes level l and, when a coarser level exists, its link to level
 * l+1; the names follow the reference (and the functions above, which this call strings together). */
typedef struct FembHierarchyLevel {
  int n_vert, n_hex;        /* dims[l], number of hexahedra                                                            */
  float spacing;            /* lattice spacing the reference assigns to the level (dx, 2 dx, 4 dx, 4 dx, ...)         */
  int* hex;                 /* [n_hex*8]  hexs[l]                                                                      */
  float* x;                 /* [n_vert*3] lattice positions (object.py:230)                                            */
  int* v2i;                 /* [n_vert]   vertex2index[l]                                                              */
  int* i2v;                 /* [n_vert]   index2vertex[l]                                                              */
  int color_vertex_num[9];
  int L_nnz;                /* blocks in the strictly lower (= upper) part; nnz = 2 L_nnz + n_vert                     */
  int* L_ptr;               /* [n_vert+1] */
  int* U_ptr;               /* [n_vert+1] */
  int* L_col;               /* [L_nnz]    */
  int* U_col;               /* [L_nnz]    */
  int* elem2nnz;            /* [n_hex*64] hex_update_offset                                                            */
  int* block_offset;        /* [2 L_nnz + n_vert] UtAUs_block_offset                                                   */
  /* link to level l+1 (all NULL / 0 on the coarsest level) */
  int* hex_parent;          /* [n_hex] coarse hexahedron of every hexahedron                                           */
  int* fine_coarse;         /* [level[l+1].n_hex*27]                                                                   */
  int n_entries;            /* interpolation entries                                                                   */
  int* coarse_id;           /* [n_entries] rowInd_noOrder                                                              */
  int* fine_id;             /* [n_entries] colInd_noOrder                                                              */
  unsigned char* entry_k;   /* [n_entries]                                                                             */
  float* weight;            /* [n_entries] un-normalised weights                                                       */
  float* weight_norm;       /* [n_entries] weight / norm[coarse_id]                                                    */
  float* norm;              /* [level[l+1].n_vert]                                                                     */
  int* ut_off;              /* [level[l+1].n_vert+1] CSR of Ut / Ut_hat in index order (coarse x fine)                 */
  int* ut_col;              /* [n_entries] */
  float* ut_w;              /* [n_entries] Ut     */
  float* ut_hat_w;          /* [n_entries] Ut_hat */
  int* us_off;              /* [n_vert+1] CSR of Us / Us_hat in index order (fine x coarse)                            */
  int* us_col;              /* [n_entries] */
  float* us_w;              /* [n_entries] Us     */
  float* us_hat_w;          /* [n_entries] Us_hat */
} FembHierarchyLevel;

typedef struct FembHierarchy {
  int n_levels;
  float origin[3];          /* lattice origin used (the bounding-box minimum unless the caller gave one)               */
  FembHierarchyLevel* level;
} FembHierarchy;

/* n_levels >= 1 is the caller's: the reference derives it from the bounding box in float32 torch arithmetic
 * (object.py:140-143; hierarchy.level_count restates it).  corrected_spacing != 0 coarsens level l with spacing dx 2^l
 * instead of the reference's 2 dx for every l >= 1 (object.py:175-179).  origin may be NULL. */
int femb_setup_hierarchy_host(int n_vert, int n_hex, const float* x, const int* hex, float dx, int n_levels, int corrected_spacing,
                              const float* origin, FembHierarchy** out);
int femb_setup_hierarchy_free(FembHierarchy* h);
/* symbolic plan of A_c = Ut_hat (A_f Us_hat) (FembGalerkinPlan): t_ptr [n_fine+1] first, then the arrays sized from it:
 * t_col [t_nnz], c_ptr [n_coarse+1], c_col / c_slot [nnz_c], t_slot [nnz_f*8], c_slot8 [t_nnz*8], ut_q [nnz(Ut_hat)];
 * fails if the product pattern is not the coarse hexahedral adjacency (block_values_reorder relies on it, object.py:1261) */
int femb_setup_galerkin_count_host(int n_fine, const int* fL_ptr, const int* fL_col, const int* fU_ptr, const int* fU_col,
                                   const int* us_off, const int* us_col, int* t_ptr);
int femb_setup_galerkin_fill_host(int n_fine, int n_coarse, const int* fL_ptr, const int* fL_col, const int* fU_ptr, const int* fU_col,
                                  const int* cL_ptr, const int* cL_col, const int* cU_ptr, const int* cU_col, const int* us_off,
                                  const int* us_col, const float* us_w, const int* ut_off, const int* ut_col, const int* t_ptr,
                                  int* t_col, int* c_ptr, int* c_col, int* c_slot, unsigned char* t_slot, unsigned char* c_slot8,
                                  unsigned char* ut_q, int* dyadic_out);

/* ---- rest state (setup) ------------------------------------------------------------------- */
/* prepare_kernal (quai_hexagon.py:4-18): rest Jacobian per (hex,q) -> detJ[n_hex*8], invJ[n_hex*72] */
int femb_prepare_rest(int n_hex, const int* hex, const float* x_rest, float* detJ, float* invJ, void* stream);
/* setup of FembLevel.fold_plan (all other fields of *lvl must be final): the inverse of the element->nnz map per row, so
 * that the assembly is a gather in a fixed order (replaces the atomics of quai_hexagon.py:100-106) */
int femb_build_fold_plan(const FembLevel* lvl, unsigned char* fold_plan, void* stream);
/* setup of FembLevel.lat_inc8 / lat_row_rec (all other fields of *lvl must be final; inc8 [n_vert*8], row_rec
 * [n_vert*FEMB_ROW_REC_BYTES]); bad (int[1], zeroed by the caller) is set if the mesh is not a lattice subset (a row with
 * more than 27 blocks, two hexahedra holding a vertex as the same local node, or a lattice offset that names two columns):
 * leave the level's lat_* fields NULL then */
int femb_build_lattice_fold(const FembLevel* lvl, int* inc8, unsigned char* row_rec, int* bad, void* stream);
/* Sum_8 + prepare_mass (cublas.py:215-218, quai_hexagon.py:20-26) as a gather: vol[n_hex], mass[n_vert] */
int femb_lumped_mass(const FembLevel* lvl, float* vol, float* mass, void* stream);

/* ---- element physics ---------------------------------------------------------------------- */
/* compute_partial_elastic_energy_X + compute_partial_gravity_energy_X + compute_partial_fixed_energy_X
 * (quai_hexagon.py:109-186; object.py:1234-1236): out = -dE/dx, gathered without atomics.
 *   x [n_vert*3] vertex order; mass [n_vert] or NULL (no gravity); pin [n_vert] 0/1 flags with per-vertex
 *   targets pin_pos [n_vert*3] (all_pin_pos) or NULL (no springs); elem_f scratch [n_hex*24];
 *   ordered != 0: out is in index order (through v2i), else vertex order (the *_noOrder kernels). */
int femb_force(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
               const int* pin, const float* pin_pos, float control_mag, int ordered, float* elem_f,
               float* out, void* stream);
/* compute_elastic_hessian (quai_hexagon.py:69-106) + the `MF_value = MF_value_fixed` copy (object.py:1233):
 * values[slot] = sum of element blocks (+ diag_add[i]*I3 on the diagonal slot of index-order row i).
 *   elem_k scratch [roundup32(n_hex)*FEMB_ELEM_K_FLOATS] (symmetric half of the element matrices); diag_add [n_vert] (index
 *   order) or NULL.
 *   If elem_f != NULL the element forces are produced by the same element pass and `force_out` is
 *   filled as in femb_force (same optional arguments). */
int femb_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* diag_add,
                 float* elem_k, float* values, void* stream);
/* tuning of the row folds: consecutive rows per warp of the fold-plan kernel (default 4) and the occupancy target of both fold
 * kernels (CTAs per SM; fold plan: 4 or 5, lattice fold: 3, 4 or 5; default 4) */
int femb_set_fold_config(int rows_per_warp, int min_blocks_per_sm);
/* lattice levels, asynchronous data movement of the assembly stage: mode 0 = none (per-lane loads, store loop); 1 = the
 * element kernel's 360 x 32 tile leaves as three TMA tensor stores (needs the driver's cuTensorMapEncodeTiled, fetched at
 * run time; without it mode 0 runs); 2 = additionally the rest data (invJ, detJ) of a CTA's hexahedra arrive through bulk
 * copies into shared memory; 3 (default) = additionally every CTA of the element kernel / the fold prefetches the index
 * data (vertex indices / incidence lists and row records) of a tile a quarter wave ahead into L2.  Results are
 * bit-identical in every mode. */
int femb_set_assembly_async(int mode);
int femb_force_hessian(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
                       const int* pin, const float* pin_pos, float control_mag, const float* diag_add,
                       float* elem_f, float* elem_k, float* force_out, float* values, void* stream);
/* compute_elastic_energy + compute_gravity_energy (quai_hexagon.py:28-64): energy[0] = total;
 * scratch [>= 1024] floats. */
int femb_energy(const FembLevel* lvl, const float* x, float mu, float lambda, const float* mass, float g,
                const int* pin, const float* pin_pos, float control_mag, float* scratch, float* energy,
                void* stream);
/* compute_fix_hessian (cublas.py:234-242): coarse pin penalty onto coarse diagonal slots.
 * Deterministic gather over the coarse rows of `ut_no_order_hat` (vertex-order, un-normalised weights,
 * the same (coarse, w) pairs as fix_idx/fix_values): values[off_d + v2i_c[J]] += k * sum_{pinned i} w_Ji^2 * I3 */
int femb_coarse_pin_penalty(const FembLevel* coarse, const FembTransfer* ut_no_order_hat, const int* pin_fine,
                            float control_mag, float* values, void* stream);

/* ---- sparse algebra on the fixed [L;D;U] pattern ------------------------------------------ */
/* bsr_mv (sparse.py:1143-1261) on self.L/self.D/self.U/self.A: y = alpha * (parts of A) x + beta * y */
int femb_spmv(const FembLevel* lvl, const float* values, int parts, float alpha, const float* x, float beta,
              float* y, void* stream);
/* r = b - A x with all three parts (downSample / showError, object.py:1022-1026,1057-1059) */
int femb_residual(const FembLevel* lvl, const float* values, const float* b, const float* x, float* r, void* stream);
/* PerformGaussSeidel (object.py:886-929): `iterations` x (backward + forward 8-colour block sweep),
 * dx starts at 0, diagonal blocks "inverted" by solve3x3 (cublas.py:59-91).  not_spd (int[1], may be
 * NULL) counts the `print('not spd')` events (cublas.py:79-80). */
int femb_gauss_seidel(const FembLevel* lvl, const float* values, const float* b, float* dx, int iterations,
                      int* not_spd, void* stream);
/* The colour launches of a sweep are programmatic dependent launches (each colour stages its matrix rows while the
 * previous colour drains); femb_set_pdl(0) makes them ordinary launches (tests).  Returns the old setting. */
int femb_set_pdl(int enable);
/* threads per row of the colour kernel: 8 (default, measured faster with 16-row blocks) or 4; the summation tree differs, so results differ in rounding */
int femb_set_gs_threads_per_row(int subs);
/* rows per CTA of the colour kernel: 16 (default: more independent row blocks in flight per SM) or 32 */
int femb_set_gs_rows_per_cta(int rows);
/* A run of colour launches of the in-place sweep above: colours first_color..last_color (descending if
 * first > last), dx is NOT zeroed.  femb_gauss_seidel(it=1) == zero dx; colours 7..0 with parts=U; colours 0..7
 * with parts=L|U.  Lets a caller put a halo exchange between colour groups (domain decomposition). */
int femb_gs_colors(const FembLevel* lvl, const float* values, const float* b, float* dx, int first_color,
                   int last_color, int parts, int* not_spd, void* stream);
/* PerformJacobi (object.py:869-884); scratch [n_vert*3] */
int femb_jacobi(const FembLevel* lvl, const float* values, const float* b, float* dx, float* scratch,
                int iterations, int* not_spd, void* stream);
/* PerformConjugateGradient / _WithInitX (object.py:931-1019) without host read-backs: the scalar
 * recurrences and the three exit tests (r1 < EPSILON, r1 <= tol*r, dot < 1e-10, k > iterations) are
 * evaluated on the device; work = r[n*3], p[n*3], ap[n*3] ; state = float[64] scratch.
 * with_init_x != 0 keeps dx as the initial guess. */
int femb_cg(const FembLevel* lvl, const float* values, const float* b, float* dx, float* r, float* p, float* ap,
            float* state, int iterations, float tol, int with_init_x, void* stream);
/* spd_matrix33f (cublas.py:115-125) on n blocks in place */
int femb_spd_blocks(float* values, int n_blocks, float spd_value, void* stream);
/* scalar-weight CSR product y = alpha * T x + beta * y (bsr_mv on Ut/Us/..., e.g. object.py:1027,1032) */
int femb_transfer_mv(const FembTransfer* t, float alpha, const float* x, float beta, float* y, void* stream);
/* A_c = Ut_hat (A_f Us_hat), spd_matrix33f on every block (if spd != 0), written in [L;D;U] layout
 * (object.py:1259-1261).  t_values scratch [t_nnz*9]. */
int femb_galerkin(const FembLevel* fine, const FembLevel* coarse, const float* fine_values,
                  const FembTransfer* ut_hat, const FembTransfer* us_hat, const FembGalerkinPlan* plan,
                  float* t_values, int spd, float spd_value, float* coarse_values, void* stream);

/* ---- dynamic (implicit Euler) step: exp1/simulation/hexagon.py:174-203, cublas.py:234-299, object.py:1286-1325 ---- */
/* Fixed_Update_Kernel + x_old = x + Basic_Update_Kernel (v *= damping; x += v dt) + x_inertia = x; dir_host: float[3] on the HOST */
int femb_dyn_begin(int n_vert, float* x, float* v, float* old_x, float* inertia_x, float* fixed_x, const int* more_fixed,
                   const int* pin, const float* dir_host, float damping, float dt, void* stream);
/* Hessian_Diag_Kernel + Hessian_Mass_Kernel as the diag_add vector of femb_hessian: diag[v2i[v]] = (pinned|dragged ? k : 0) + m/dt^2 */
int femb_dyn_diag(const FembLevel* lvl, const float* mass, const int* more_fixed, const int* pin, float control_mag,
                  float inv_dt, float* diag, void* stream);
/* compute_partial_more_fixed_energy_X + compute_Inertia_Gradient_Kernel onto an index-order force */
int femb_dyn_force(const FembLevel* lvl, const float* x, const float* inertia_x, const int* more_fixed, const float* fixed_x,
                   const float* mass, float control_mag, float inv_dt, float* grad, void* stream);
/* updateVelocity: v = (x - x_old) * inv_dt */
int femb_dyn_velocity(int n_vert, const float* x, const float* old_x, float inv_dt, float* v, void* stream);
/* Control_Kernel: more_fixed[i] = |x_i - x_select|^2 < 0.002 */
int femb_control(int n_vert, int select_v, const float* x, int* more_fixed, void* stream);

/* ---- mass-spring cloth: exp1/cloth_simulation/triangle.py:5-79, object.py:657-677 there ------------------------- */
/* compute_rest_length: rest[e] = |x[p0] - x[p1]|, edge = int[n_edge*2] */
int femb_cloth_rest_length(int n_edge, const int* edge, const float* x, float* rest, void* stream);
/* Hessian_Mass_Kernel + compute_elastic_hessian + compute_elastic_gradient + gravity + inertia, without atomics.
 * inc_ptr/inc: CSR vertex -> (edge*2 + end) incidence; off_p0p1/off_p1p0/d_slot: value slots of the blocks (p0,p1),
 * (p1,p0) of every edge and of every diagonal block (the reference's `offset`/`d_offset` maps, object.py:139-176 there,
 * translated to the caller's value layout); edge_h [n_edge*9], edge_g [n_edge*3] scratch; grad in vertex order. */
int femb_cloth_assemble(int n_vert, int n_edge, const int* edge, const float* rest, float k, const float* x, const float* mass,
                        float g, const float* inertia_x, float inv_dt, const int* inc_ptr, const int* inc,
                        const int* off_p0p1, const int* off_p1p0, const int* d_slot, float* edge_h, float* edge_g,
                        float* values, float* grad, void* stream);
/* the cloth's update_deltaX_kernel (cublas.py:22-26 there): x += dx except vertices skip0, skip1 (0 and 65) */
int femb_cloth_update(int n_vert, const float* dx, float* x, int skip0, int skip1, void* stream);

/* ---- whole steps (csrc/femb_steps.cu): one call = one loop body of a solver method of `class Object` ---------------------
 * The context structs are HOST structs of device pointers (the arrays the reference allocates in Object.__init__,
 * object.py:797-855); arrays indexed by level are host arrays of `layers` (or `layers - 1`) entries.  Every call only
 * enqueues on `stream`: capture it into a CUDA graph and replay it per step. */
typedef struct FembContext {
  int layers;                        /* self.layer                                                                    */
  const FembLevel* levels;           /* [layers]                                                                      */
  const FembTransfer* Ut_hat;        /* [layers-1] residual restriction  (index order, un-normalised; object.py:1027)  */
  const FembTransfer* Us_hat;        /* [layers-1] prolongation          (index order; object.py:1032)                */
  const FembTransfer* Ut_noOrder;    /* [layers-1] state restriction     (vertex order, row-normalised; :1489) or NULL */
  const FembTransfer* Ut_noOrder_hat;/* [layers-1] (vertex order, un-normalised; coarse pin penalty :1501) or NULL     */
  const FembGalerkinPlan* plans;     /* [layers-1] or NULL (steps without Galerkin products)                           */
  float* const* values;              /* [layers] MF_value_gpu, UtAUs_values[0..]                                      */
  float* const* t_values;            /* [layers-1] scratch of femb_galerkin (plans[l].t_nnz * 9 floats) or NULL        */
  float* const* B;                   /* [layers] dev_B                                                                */
  float* const* B_fixed;             /* [layers] dev_B_fixed                                                          */
  float* const* delta_x;             /* [layers] dev_delta_x                                                          */
  float* const* x_solved;            /* [layers] dev_x_solved                                                         */
  float* const* R;                   /* [layers] dev_R                                                                */
  float* x;                          /* x_gpu_layer[0]: the state, vertex order                                       */
  float* grad;                       /* grad_gpu                                                                      */
  const float* diag_fixed;           /* the k*I pin blocks of MF_value_fixed as a per-row diagonal (index order)      */
  const float* mass0;                /* m_gpu[0]                                                                      */
  const int* pin;                    /* pin_gpu                                                                       */
  const float* pin_pos;              /* all_pin_pos_gpu                                                               */
  float* elem_f;                     /* scratch [max n_hex * 24]                                                      */
  float* elem_k;                     /* scratch [max n_hex * FEMB_ELEM_K_FLOATS]                                      */
  int* not_spd;                      /* int[1] counter of solve3x3 'not spd' events, or NULL                          */
  float mu, la, g, control_mag, spd_value;
  int spd_level0;                    /* exp1/simulation applies spd_matrix33f on level 0 as well                      */
} FembContext;
/* VCycle(layer) (object.py:1216-1226): 3 pre + 3 post symmetric GS per level, 3 on the coarsest; reads dev_B_fixed[layer] */
int femb_vcycle(const FembContext* ctx, int layer, void* stream);
/* the Galerkin chain A_i = Ut_hat (A_{i-1} Us_hat) + spd_matrix33f for i = 1 .. layers-1 (object.py:1252-1264) */
int femb_galerkin_chain(const FembContext* ctx, void* stream);
/* the loop body of NewtonMultigrid (object.py:1231-1296): force + Hessian, Galerkin chain, one V-cycle, finish(), x update */
int femb_newton_mg_step(const FembContext* ctx, void* stream);
/* the loop body of FASv1 / FASv2 / FASv3 (object.py:1373-1511; variant 1..3) on the state `x` (NULL = ctx->x; FASv1 iterates
 * on dev_x_solved[0]) */
int femb_fas_step(const FembContext* ctx, int variant, float* x, void* stream);

/* dynamic (implicit Euler) frame: exp1/simulation/object.py:1286-1325 = one Newton iteration with 5 CG iterations */
typedef struct FembDynamic {
  float* x;               /* dev_x (the state the frame advances)                  */
  float* v;               /* dev_v                                                 */
  float* old_x;           /* dev_old_x                                             */
  float* inertia_x;       /* dev_inertia_x                                         */
  int* more_fixed;        /* more_fixed_gpu (drag constraint flags)                */
  float* diag_dyn;        /* scratch [n_vert]: pin / drag / mass diagonal          */
  float* cg_p;            /* dev_P[0]                                              */
  float* cg_ap;           /* dev_AP[0]                                             */
  float* cg_state;        /* float[>= 1040] scratch of femb_cg                     */
  const float* dir_host;  /* float[3] on the HOST: drag direction of this frame    */
  float dt, inv_dt, damping;
} FembDynamic;
int femb_dynamic_step(const FembContext* ctx, const FembDynamic* dyn, void* stream);

/* cloth frame: exp1/cloth_simulation/object.py:657-677 (arguments as in femb_cloth_assemble / femb_cg) */
typedef struct FembCloth {
  const FembLevel* level;      /* pattern of the cloth matrix ([L;D;U] view, identity ordering)                    */
  int n_edge;
  const int* edge; const float* rest; const float* mass;
  const int* inc_ptr; const int* inc; const int* off_p0p1; const int* off_p1p0; const int* d_slot;
  const int* zeros_i;          /* int[n_vert] of zeros (no pins / drags in the cloth)                              */
  float* x; float* v; float* old_x; float* inertia_x;
  float* values; float* grad; float* B; float* B_fixed; float* delta_x; float* cg_p; float* cg_ap; float* cg_state;
  float* edge_h; float* edge_g;
  float k, g, dt, inv_dt, damping;
} FembCloth;
int femb_cloth_step(const FembCloth* cloth, void* stream);

/* ---- multi-GPU: communicator, halo exchange, scalar all-reduce (csrc/femb_comm.cu) -------------------------------
 * The reference is single-GPU (no NCCL / MPI call site anywhere); a mesh-partitioned caller of this ABI
 * (fem-simulation_b200/distributed.py) needs exactly these.  One FembComm per rank = one NCCL communicator on the
 * calling thread's current device; NCCL is loaded at femb_comm_init (dlopen("libnccl.so.2")), the library does not
 * link against it.  All calls only enqueue on `stream` (graph-capturable once the communicator's connections exist,
 * i.e. after one eager exchange with every neighbour). */
#define FEMB_COMM_ID_BYTES 128          /* sizeof(ncclUniqueId) */
#define FEMB_MAX_NEIGHBOURS 8
#define FEMB_REDUCE_SUM 0
#define FEMB_REDUCE_MAX 1
typedef struct FembComm FembComm;
/* Ghost-row exchange plan of one vector layout: for neighbour k (rank nb[k]) the rows this rank sends and the ghost
 * rows it receives; the lists of all neighbours are concatenated in neighbour order.  Both sides order a message by
 * the lattice position of its rows.  send_buf / recv_buf: device staging, (sum n_send) * width and (sum n_recv) * width
 * floats. */
typedef struct FembHalo {
  int n_nb;
  int nb[FEMB_MAX_NEIGHBOURS];
  int n_send[FEMB_MAX_NEIGHBOURS];
  int n_recv[FEMB_MAX_NEIGHBOURS];
  const int* send_rows;
  const int* recv_rows;
  float* send_buf;
  float* recv_buf;
} FembHalo;
/* id_host: FEMB_COMM_ID_BYTES bytes on the host; created on one rank, handed to the others by the caller (any channel) */
int femb_comm_unique_id(unsigned char* id_host);
int femb_comm_init(FembComm** comm, const unsigned char* id_host, int rank, int nranks);
/* with the peer-memory transport: call on every rank at the same point, after a barrier of the caller's (no rank may still
 * be storing into another's arena) */
int femb_comm_destroy(FembComm* comm);
/* Peer-memory transport for femb_halo_exchange (one NVLink box, one process per GPU).  COLLECTIVE over the communicator:
 * nb_host[n_nb] = every rank this rank will ever exchange halos with (the relation must be symmetric), slot_bytes >= the
 * largest per-neighbour segment (rows * width * 4) of any plan.  Allocates the receive arena of this rank (2 * FEMB_MAX_NEIGHBOURS
 * slots + flags), hands its CUDA-IPC handle to the other ranks and maps theirs.  If any rank cannot map all its neighbours the
 * call fails on EVERY rank and the communicator keeps using NCCL send / recv. */
int femb_comm_p2p_enable(FembComm* comm, const int* nb_host, int n_nb, size_t slot_bytes, void* stream);
/* err_host[0] = 1 if a pull kernel gave up waiting for a neighbour's data (~30 s); synchronises the device */
int femb_comm_p2p_error(FembComm* comm, int* err_host);
/* vec[recv rows] <- the neighbours' vec[send rows]; vec is float[n * width] (width 3: a level vector, 1: a scalar field), all on
 * `stream`.  Peer memory enabled: ONE kernel (gather + NVLink stores into the neighbours' arenas + epoch flag, then wait for the
 * neighbours' flags and scatter into the ghost rows); send_buf / recv_buf are not used.  Otherwise: pack kernel, one NCCL
 * group of send / recv pairs, unpack kernel.  Every rank must issue the same sequence of exchanges. */
int femb_halo_exchange(FembComm* comm, const FembHalo* halo, float* vec, int width, void* stream);
/* in-place all-reduce of n floats (batch the scalars of a step into one call): FEMB_REDUCE_SUM or FEMB_REDUCE_MAX */
int femb_comm_allreduce(FembComm* comm, float* buf, int n, int op, void* stream);

/* ---- exp2 train(): adjoint of the two-level step w.r.t. the interpolation weights (csrc/femb_exp2.cu) --------------------
 * exp2/p_hat/quasi_simulation/object.py:1202-1287 (wp.Tape) -- the forward pass uses the operators above; these are the
 * hand-written vector-Jacobian products of the taped kernels (forward mode on dual numbers, one thread per input direction). */
/* adjoint of `loss` (atomic_max of |component|, cublas.py:274-281 there): onehot[i] = sign(v[i]) where |v[i]| == maxv[0], else 0 */
int femb_absmax_onehot(int n, const float* v, const float* maxv, float* onehot, void* stream);
/* adjoint of jacobi_iteration (cublas.py:99-102 there; x = solve3x3(D, b, 0)) w.r.t. the diagonal blocks:
 * dbar[row*9+k] = lam[row] . d x[row] / d D[row][k]; diag_blocks / dbar [n*9], b / lam [n*3] (any common row order) */
int femb_solve3x3_vjp(int n, const float* diag_blocks, const float* b, const float* lam, float* dbar, void* stream);
/* adjoint of prepare_kernal + compute_mat_F + compute_elastic_hessian (diagonal blocks) of level `lvl` w.r.t. the node
 * positions x and the rest positions x_rest (both [n_vert*3], vertex order): dbar_vertex [n_vert*9] = cotangent of the diagonal
 * block of every vertex, elem_scratch [n_hex*48]; out: lam_x, lam_rest [n_vert*3] */
int femb_coarse_diag_vjp(const FembLevel* lvl, const float* x, const float* x_rest, float mu, float lambda, const float* dbar_vertex,
                         float* elem_scratch, float* lam_x, float* lam_rest, void* stream);
/* adjoint of Kronecker_product + the two bsr_mv with Ut_noOrder (x1 = Ut X, x_rest1 = Ut x_init) + loss_norm
 * (cublas.py:284-292 there): grad[p] of entry p = (J, i): lam_x[J].X[i] + lam_rest[J].x_init[i] + 2 alpha (1 - rowsum_J);
 * reg_rows [nrow] (or NULL) receives (1 - rowsum_J)^2 of the rows that count (|1 - rowsum| >= 1e-6) */
int femb_exp2_weight_grad(const FembTransfer* ut, const float* w, const float* lam_x, const float* X, const float* lam_rest,
                          const float* x_init, float alpha, float* grad, float* reg_rows, void* stream);
/* Valueaxpy (cublas.py:156-166 there): w = clamp(w + lr * grad, 0, 1) */
int femb_exp2_weight_update(int n, float* w, const float* grad, float lr, void* stream);

/* ---- vectors ------------------------------------------------------------------------------- */
int femb_axpby(int n, float a, const float* x, float b, float* y, void* stream);          /* y = a x + b y (cublas.py:160-163) */
int femb_copy(int n, const float* x, float* y, void* stream);
/* one step of the reference's Adam loop (updateM, updateV, scal_ x2, updateX: cublas.py:172-194, object.py:1090-1094) on n
 * floats: m = (1-b1) g + b1 m; v = (1-b2) g*g + b2 v; m_hat = bias1 m; v_hat = bias2 v; x -= lr m_hat / (sqrt(v_hat) + eps)
 * (bias1 = 1/(1-b1^t), bias2 = 1/(1-b2^t); the sign is the reference's: `g` is the force, cublas.py:192) */
int femb_adam_update(int n, const float* grad, float* m, float* v, float* m_hat, float* v_hat, float* x, float beta1, float beta2,
                     float bias1, float bias2, float lr, float epsilon, void* stream);
int femb_fill(int n, float v, float* y, void* stream);
/* update_deltaX_kernel (cublas.py:20-24): x[i2v[i]] += dx[i] */
int femb_update_x(const FembLevel* lvl, const float* dx, float* x, void* stream);
/* V2I / I2V (cublas.py:203-213): ordered[v2i[v]] = vert[v]  /  vert[i2v[i]] = ordered[i] */
int femb_v2i(const FembLevel* lvl, const float* vert, float* ordered, void* stream);
int femb_i2v(const FembLevel* lvl, const float* ordered, float* vert, void* stream);
/* Inf_norm (cublas.py:38-48): out[0] = max |x_k| (out is zeroed by the call) */
int femb_inf_norm(int n, const float* x, float* out, void* stream);
/* cublasSdot / square_sum (cublas.py:32-36,198-201), deterministic two-level reduction; scratch >= 1024 */
int femb_dot(int n, const float* x, const float* y, float* scratch, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEMB200_H */
