/* xgpu.h — C ABI of libxgpu.so: the B200 (sm_100a) cut-element assembly path behind ngsxfem's
 * CutInfo / dCut / SymbolicCutBFI / dFacetPatch / RestrictedBilinearForm.Assemble surface.
 *
 * Conventions
 *   - plain C, no torch / NGSolve types.  Every array pointer is a DEVICE pointer unless the
 *     parameter name ends in `_host`.  All arrays are caller owned; the library allocates only
 *     its private scratch (inside xgpu_ctx / xgpu_plan).
 *   - every call returns 0 (XGPU_OK) or a negative error code and never throws; the message of
 *     the last error is available through xgpu_last_error().  Geometry errors that the
 *     reference reports by throwing per element (straightcutrule.cpp:45-49,144-149,190-194)
 *     are recorded on the device and fetched with xgpu_get_status() (first failing element).
 *   - work is enqueued on the context's stream (xgpu_set_stream, a cudaStream_t); calls whose
 *     result is a host scalar (`*_host` out-parameters) synchronise that stream.
 *   - reference element conventions are those pinned by cutint/straightcutrule.cpp:506-514,
 *     straightcutrule.hpp:56-60: TET v0=(1,0,0) v1=(0,1,0) v2=(0,0,1) v3=(0,0,0) etc.
 *   - bitsets are arrays of uint32 words, bit i = word i>>5, bit i&31 — byte-identical to
 *     NGSolve's BitArray (bit i = byte i>>3, bit i&7) on little-endian hosts.
 *
 * Each entry point cites the reference interface it replaces (paths relative to ngsxfem).
 */
#ifndef XGPU_H
#define XGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* utils/ngsxstd.hpp:16-19 */
enum { XGPU_NEG = 0, XGPU_POS = 1, XGPU_IF = 2 };

enum {
  XGPU_OK = 0,
  XGPU_ERR_CUDA = -1,
  XGPU_ERR_ARG = -2,
  XGPU_ERR_UNSUPPORTED = -3,   /* the reference's legacy paths (subdivlvl>0, quads/hexes, ...) */
  XGPU_ERR_GEOM_PLAIN = -11,   /* straightcutrule.cpp:45-49 (zero vertex value on a cut element) */
  XGPU_ERR_GEOM_SCUT = -12,    /* straightcutrule.cpp:144-149 */
  XGPU_ERR_GEOM_NOTIMPL = -13, /* straightcutrule.cpp:190-194 */
  XGPU_ERR_CAPACITY = -14
};

typedef struct xgpu_ctx xgpu_ctx;
typedef struct xgpu_plan xgpu_plan;

/* ---------------------------------------------------------------- context */
int xgpu_create(int device, xgpu_ctx** ctx_out);
void xgpu_destroy(xgpu_ctx* ctx);
int xgpu_set_stream(xgpu_ctx* ctx, void* cuda_stream);
/* tuning switches (no reference counterpart).  "elmat_dmma": element-matrix contraction on FP64
 * tensor-core tiles: -1 never, 0 (default) at FE order >= 3, 1 also at order 2.  "short_blocks": cut quadrature / cut element matrices with one tile per block
 * instead of a persistent grid (for a low-priority side stream; default 0).  "csr_small": one-pass CSR, rows with few adjacent table rows on the thread-per-row
 * kernels (default 1).  "csr_warp": one-pass CSR, the other rows with at most 256 candidates on the warp-per-row kernels (default 1).
 * "csr_warp_half": warp-per-row pattern pass, rows with at most 128 candidates on the 4-keys-per-lane sorting network (default 1; same records).
 * "csr_trips": one-pass CSR lane-group kernels, 2 = half-width lane groups with two columns per
 * lane (0 = policy).  "csr_closed_form": one-pass CSR, rows of procedural affine P2-tet batches from the closed-form barycentric
 * expressions instead of the reference tensors (default 1; equal up to rounding).  "csr_fork": one-pass CSR, the warp-per-row kernel runs on an internal second stream next to the
 * thread-per-row kernel (default 1; they take disjoint rows).  "elmat_lp": cut-element matrices at order <= 2 on the lane-per-point kernel
 * (default 1; 0 = the staged half-warp kernel).  "elmat_st_tp": space-time element
 * matrices on the tensor-structured kernel where instantiated (default 1; 0 = always the generic kernel) */
int xgpu_set_option(xgpu_ctx* ctx, const char* key, int value);
/* counters of a context (the reference has RegionTimer / static Timer objects, xfem/symboliccutbfi.cpp:88,495):
 * kernels launched, instrumented entry points called (each is an NVTX range "xgpu_<name>"), plans created, bytes of the
 * grow-only scratch arena, and the sticky geometry status (code + first failing element, as xgpu_get_status).  Synchronises. */
typedef struct {
  int64_t kernel_launches, api_calls, plans_created, scratch_bytes;
  int status_code;
  int64_t status_element;
} xgpu_stats;
int xgpu_get_stats(xgpu_ctx* ctx, xgpu_stats* out);
/* Isoparametric mesh deformation — the state NGSolve keeps in the mesh (`mesh.SetDeformation(gf)`, set and
 * unset around every assembly by the `deformation=` argument of dCut / dFacetPatch / Integrate,
 * python/__init__.py:361-439,749-864; cutint/cutintegral.cpp:126-212): while set, every geometric call
 * below (xgpu_cutquad_fill for dt=IF, xgpu_elmat_*, xgpu_facetpatch, xgpu_integrate_monomial) uses
 *   x(xhat) = x_affine(xhat) + sum_j d_j phi_j(xhat),   F = F_affine + sum_j d_j (x) grad phi_j
 * with d in H1(order)^dim: el_dofs ne x xgpu_h1_ndof(dim, order) (same mesh element numbering as el_vtx),
 * vals ndof x dim.  The arrays must stay valid until the deformation is unset (order = 0 / NULL).
 * Multi-level-set and space-time rules do not support a deformation (XGPU_ERR_UNSUPPORTED). */
int xgpu_set_deformation(xgpu_ctx* ctx, int order, const int32_t* el_dofs, const double* vals);
int xgpu_sync(xgpu_ctx* ctx);
const char* xgpu_last_error(xgpu_ctx* ctx);
/* per-element geometry error recorded by kernels since the last xgpu_clear_status; syncs. */
int xgpu_get_status(xgpu_ctx* ctx, int* code_host, int64_t* elem_host);
int xgpu_clear_status(xgpu_ctx* ctx);
/* number of kernels this context has launched so far (bench: gpu_launches) */
int64_t xgpu_launch_count(xgpu_ctx* ctx);
const char* xgpu_version(void);

/* ---------------------------------------------------------------- tables (host side)
 * NGSolve's SelectIntegrationRule / H1HighOrderFE are third party; defaults restate their
 * published construction, xgpu_set_rule injects the exact tables when an adapter has them
 * (call sites: straightcutrule.cpp:87-90,673; symboliccutbfi.cpp:2009-2012). */
int xgpu_rule_size(xgpu_ctx* ctx, int dim, int order);
int xgpu_get_rule(xgpu_ctx* ctx, int dim, int order, double* pts_host /* n x 3 */, double* wts_host);
int xgpu_set_rule(xgpu_ctx* ctx, int dim, int order, int n, const double* pts_host, const double* wts_host);
int xgpu_h1_ndof(int dim, int order);
/* host evaluation of the device basis (tests / adapters): shape[nd], dshape[nd x dim] */
int xgpu_h1_shapes_host(int dim, int order, const int32_t* vnum_host, const double* xhat_host,
                        double* shape_host, double* dshape_host);

/* ---------------------------------------------------------------- (a) CutInfo
 * CutInformation::Update, VOL/BND loop (xfem/cutinfo.cpp:61-133): classification of simplices
 * of dimension `dim` from nodal P1 level-set values.  cut_ratio may be NULL. */
int xgpu_classify(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phi,
                  uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if, double* cut_ratio);
/* CutInformation::Update with a space-time level set (time_order >= 0, cutinfo.cpp:90-124): the part volumes are the
 * weight sums of the order-0 NEG and POS rules (xgpu_cutquad_st_*; added in rule order), the element is IF iff both are
 * > 0, NEG iff only the NEG one, POS otherwise; cut_ratio = neg / (neg + pos) (nullable). */
/* the same decision in ONE pass over the elements (no rule storage): part volumes of the order-0 POS / NEG space-time
 * rules summed in rule order (bit-identical to xgpu_cutquad_st_* + xgpu_rule_element_sums), ballot classification.
 * phi_st: (time_fe_order+1) x nv nodal values, time-major. */
int xgpu_classify_st(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phi_st, int64_t nv,
                     int time_fe_order, int time_order, uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if,
                     double* cut_ratio);
int xgpu_rule_element_sums(xgpu_ctx* ctx, int64_t n, const int64_t* offsets, const double* wts, double* sums);
int xgpu_classify_volumes(xgpu_ctx* ctx, int64_t ne, const double* vol_neg, const double* vol_pos,
                          uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if, double* cut_ratio);
/* BitArray algebra used by cutinfo.cpp:127-132 (UNCUT/HASNEG/HASPOS) and by user code.
 * op: 0 OR, 1 AND, 2 ANDNOT (a & ~b), 3 NOT a, 4 COPY a, 5 CLEAR, 6 SET (tail bits stay 0) */
int xgpu_bitset_op(xgpu_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int64_t nbits);
int xgpu_bitset_count(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int64_t* count_host);
/* counts of k <= 8 bitsets with ONE host wait (bits_host: host array of device pointers): counts_host[i] = set bits of
 * bitset i with index < nbits_host[i] — a prefix of a longer bitset may be counted (own elements before halo elements) */
int xgpu_bitset_count_many(xgpu_ctx* ctx, int k, const uint32_t* const* bits_host, const int64_t* nbits_host,
                           int64_t* counts_host);
/* ascending list of set bits; list must hold xgpu_bitset_count entries */
int xgpu_bitset_to_list(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int32_t* list);
/* CutInformation::GetElementsWithThresholdContribution (cutinfo.cpp:263-284) */
int xgpu_threshold_elements(xgpu_ctx* ctx, int64_t ne, const double* cut_ratio, int dt, double threshold,
                            uint32_t* bits_out);
/* MultiLevelsetCutInformation::UpdateElementsOfDomainType (cutinfo.cpp:356-398).
 * phis: M x nv (row j = level set j); dts_host: ntuples x M */
int xgpu_classify_ml(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phis,
                     int64_t nv, int M, const int* dts_host, int ntuples, uint32_t* bits_out);
/* GetFacetsWithNeighborTypes (cutinfo.cpp:476-538); facet_els: nf x 2, -1 = boundary */
int xgpu_facets_with_neighbor_types(xgpu_ctx* ctx, int64_t nf, const int32_t* facet_els,
                                    const uint32_t* a, const uint32_t* b, int bnd_val_a, int bnd_val_b,
                                    int use_and, uint32_t* bits_out);
/* GetElementsWithNeighborFacets (cutinfo.cpp:540-566) */
int xgpu_elements_with_neighbor_facets(xgpu_ctx* ctx, int64_t nf, const int32_t* facet_els,
                                       const uint32_t* facet_bits, int64_t ne, uint32_t* bits_out);
/* GetDofsOfElements (cutinfo.cpp:599-622); el_dofs: ne x nd, -1 = no dof */
int xgpu_dofs_of_elements(xgpu_ctx* ctx, int64_t ne, int nd, const int32_t* el_dofs,
                          const uint32_t* el_bits, int64_t ndof, uint32_t* bits_out);
/* the converse: elements with at least one marked entry in their row of `table` (ne x nd, -1 = unused).  With the vertex table
 * and xgpu_dofs_of_elements on it this is GetElementsWithSharedVertex (xfem/cutinfo.cpp:568-597): mark the vertices of the
 * elements in a, then every element touching a marked vertex. */
int xgpu_elements_with_marked(xgpu_ctx* ctx, int64_t ne, int nd, const int32_t* table, const uint32_t* marks, uint32_t* bits_out);
/* RestrictedFESpace / CompressedFESpace numbering (utils/restrictedfespace.cpp:23-51):
 * map[d] = rank of d among the set bits, -1 if unset; *nactive_host = number of set bits */
int xgpu_bitset_rank(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int32_t* map, int64_t* nactive_host);
/* out[i][k] = map ? map[table[list[i]][k]] : table[list[i]][k]   (list==NULL: identity) */
int xgpu_gather_rows(xgpu_ctx* ctx, int64_t n, const int32_t* list, int width, const int32_t* table,
                     const int32_t* map, int32_t* out);

/* ---------------------------------------------------------------- (b) cut quadrature
 * StraightCutIntegrationRule (cutint/straightcutrule.cpp:582-677) for a list of simplices
 * (elems==NULL: 0..n-1).  Points are reference coordinates of the background element
 * (3 doubles per point, unused coordinates 0); weights follow the reference convention
 * (w*|det F| is the physical measure, also for dt=IF).  Two calls: _count fills offsets[n+1]
 * (device, int64) and returns the total; _fill writes pts[total x 3], wts[total]. */
int xgpu_cutquad_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                       const double* phi, int dt, int order, int64_t* offsets, int64_t* total_host);
int xgpu_cutquad_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                      const double* vtx_xyz, const double* phi, int dt, int order,
                      const int64_t* offsets, double* pts, double* wts);
/* CreateMultiLevelsetCutIntegrationRule (cutint/mlsetintegration.cpp:6-186,
 * straightcutrule.cpp:679-812): phis M x nv, dts_host ntuples x M */
int xgpu_cutquad_ml_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                          const double* phis, int64_t nv, int M, const int* dts_host, int ntuples,
                          int order, int64_t* offsets, int64_t* total_host);
int xgpu_cutquad_ml_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                         const double* vtx_xyz, const double* phis, int64_t nv, int M,
                         const int* dts_host, int ntuples, int order, const int64_t* offsets,
                         double* pts, double* wts);
/* SpaceTimeCutIntegrationRule (cutint/spacetimecutrule.cpp:104-244): phi_st is time-major,
 * (time_fe_order+1) x nv nodal values; tvals receives the reference time of every point. */
int xgpu_cutquad_st_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                          const double* phi_st, int64_t nv, int time_fe_order, int dt, int order_time,
                          int order_space, int64_t* offsets, int64_t* total_host);
int xgpu_cutquad_st_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                         const double* vtx_xyz, const double* phi_st, int64_t nv, int time_fe_order,
                         int dt, int order_time, int order_space, const int64_t* offsets,
                         double* pts, double* tvals, double* wts);
/* Integrate (cutint/cutintegral.cpp:126-212) of a monomial prod x_i^expo[i] over a rule batch:
 * sum_q w_q |det F| f(x_q).  result_host receives the sum (fixed-order reduction). */
int xgpu_integrate_monomial(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                            const double* vtx_xyz, const int64_t* offsets, const double* pts,
                            const double* wts, const int* expo_host, double* result_host);

/* space-time Integrate (cutint/python_cutint.cpp:21-163 with time_order >= 0; tests/pytests/test_spacetimecutrule.py):
 * sum_q w_q |det F| x^e0 y^e1 z^e2 tref^e3 over a space-time rule batch (xgpu_cutquad_st_*); expo_host has 4 entries. */
int xgpu_integrate_monomial_st(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                               const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,
                               const double* wts, const int* expo_host, double* result_host);

/* the same over BOUNDARY elements (dCut(..., vb=BND), tests/pytests/test_bndcut.py): simplices of dimension `dim`
 * embedded in R^sdim (segments in 2D / 3D, triangles in 3D), measure sqrt(det F^T F); rules from
 * xgpu_cutquad_count/_fill with the boundary elements' table and dt = NEG / POS. */
int xgpu_integrate_monomial_bnd(xgpu_ctx* ctx, int dim, int sdim, int64_t n, const int32_t* elems,
                                const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets,
                                const double* pts, const double* wts, const int* expo_host, double* result_host);

/* ---------------------------------------------------------------- (c) element matrices
 * SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixAdd (xfem/symboliccutbfi.cpp:76-486)
 * for a closed menu of integrands: sum of terms  coef * h^hpow * op_test(v) * op_trial(u),
 * ops: 0 = ID, 1 = GRAD (GRADxGRAD is the dot product), 2 = DN = grad . n with
 * n = grad(lset_p1)/|grad(lset_p1)|; h = |det F|^(1/dim).  mats: n x nd x nd, row-major,
 * rows = test, cols = trial.  accumulate!=0 ADDS into mats (CalcElementMatrixAdd).  slots (nullable):
 * matrix i is written to mats[slots[i]] instead of mats[i] (several integrators, one batch). */
enum { XGPU_OP_ID = 0, XGPU_OP_GRAD = 1, XGPU_OP_DN = 2 };
typedef struct { int op_trial, op_test; double coef; int hpow; } xgpu_term;
/* per-element rules (cut elements): offsets/pts/wts from xgpu_cutquad_* */
int xgpu_elmat_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                    const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                    const int64_t* offsets, const double* pts, const double* wts,
                    const xgpu_term* terms_host, int nterms, int accumulate, const int32_t* slots,
                    double* mats);
/* two integrators over the same element list in ONE pass — NGSolve's element loop calls CalcElementMatrixAdd of every
 * integrator on an element before it moves on (e.g. the NEG volume terms and the IF interface term of a CutFEM form
 * on the cut elements): rule A with terms A, then rule B with terms B, one matrix per element. */
int xgpu_elmat_rule_pair(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                         const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                         const int64_t* offsets_a, const double* pts_a, const double* wts_a,
                         const xgpu_term* terms_a_host, int nterms_a,
                         const int64_t* offsets_b, const double* pts_b, const double* wts_b,
                         const xgpu_term* terms_b_host, int nterms_b,
                         int accumulate, const int32_t* slots, double* mats);
/* cut rules made ON THE FLY — what T_CalcElementMatrixAdd does: it calls CreateCutIntegrationRule for the element at hand
 * and contracts at once, no rule is ever stored (xfem/symboliccutbfi.cpp:105-121; cutint/straightcutrule.cpp:131-196,
 * 503-677).  Up to two integrators over the same element list in one pass (dt_b < 0: one): domain type, quadrature order
 * and terms each; a listed element that is not cut gets the standard rule if its type equals dt_x and nothing otherwise,
 * exactly as xgpu_cutquad_* would hand it on.  Same matrices as xgpu_cutquad_count/_fill + xgpu_elmat_rule_pair (to
 * rounding), without the 32 B per quadrature point written and read back.  fe_order 1..2, ID*ID and GRAD.GRAD terms,
 * affine mesh; anything else returns XGPU_ERR_UNSUPPORTED (the caller takes the two-step path).  Geometry errors are
 * recorded in the sticky status (xgpu_get_status) like those of xgpu_cutquad_fill. */
int xgpu_elmat_cut_pair(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                        const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                        int dt_a, int order_a, const xgpu_term* terms_a_host, int nterms_a,
                        int dt_b, int order_b, const xgpu_term* terms_b_host, int nterms_b,
                        int accumulate, const int32_t* slots, double* mats);
/* uncut elements: the shared standard rule of `order` on every listed element */
int xgpu_elmat_std(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                   const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int order,
                   const xgpu_term* terms_host, int nterms, int accumulate, const int32_t* slots,
                   double* mats);

/* uncut AFFINE elements, mass / stiffness terms only (ID*ID, GRAD.GRAD): the element matrix is
 *   sum_m coefs[m] * tref[m]   with  coefs = |det F| * (c_m, c_s G_00, c_s G_01, ...),  G = F^-1 F^-T  (1 + dim(dim+1)/2 entries,
 * rows padded to 8 doubles) and the reference tensors tref[m] (nd x nd each: M_ref, K_ref^ab) of the rule of `order`.
 * xgpu_affine_coefs writes the per-element coefficients (64 B instead of nd^2 * 8 B); xgpu_reftensor returns the DEVICE
 * pointer of the cached tensors.  Consumed by procedural batches of xgpu_csr_assemble_fused. */
int xgpu_affine_coefs(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                      const double* vtx_xyz, const xgpu_term* terms_host, int nterms, double* coefs /* n x 8 */);
int xgpu_reftensor(xgpu_ctx* ctx, int dim, int fe_order, int order, const double** tref_dev_out, int* ncoef_host);

/* element vectors: SymbolicCutLinearFormIntegrator::T_CalcElementVector (xfem/symboliccutlfi.cpp:51-140) for
 * sums of  f(x) * op_test(v),  f = coef * h^hpow * x^expo[0] y^expo[1] z^expo[2],  op_test 0 = ID, 2 = DN.
 * vecs: n x nd; rule order of _std: `order`, or 2*fe_order if < 0 (:86-88).  Same accumulate / slots meaning. */
typedef struct { int op_test; double coef; int hpow; int expo[3]; } xgpu_lterm;
int xgpu_elvec_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                    const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                    const int64_t* offsets, const double* pts, const double* wts,
                    const xgpu_lterm* terms_host, int nterms, int accumulate, const int32_t* slots,
                    double* vecs);
int xgpu_elvec_std(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                   const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int order,
                   const xgpu_lterm* terms_host, int nterms, int accumulate, const int32_t* slots,
                   double* vecs);

/* space-time element matrices / vectors (BASELINE config 5): T_CalcElementMatrixAdd (xfem/symboliccutbfi.cpp:76-486)
 * on a SpaceTimeFE (spacetime/SpaceTimeFE.cpp:57-135) with the rule of xgpu_cutquad_st_* (reference point, reference
 * time, weight).  Basis: shape[j*nd_s + i] = L_j(t) N_i(xhat), time-major (SpaceTimeFE.hpp:185-218), L_j = Lagrange
 * polynomials on NodalTimeFE's nodes (SpaceTimeFE.cpp:258-282).  Operators: 0 = ID, 1 = GRAD (space gradient,
 * GRAD.GRAD only), XGPU_OP_DT = derivative w.r.t. the REFERENCE time (dt(u), spacetime/diffopDt.hpp).  fix_tref(u, t0)
 * terms are the same call with a space rule and tvals = t0.  mats: n x nd x nd, nd = nd_s*(time_fe_order+1),
 * rows = test; measure = spatial |det F|.  No deformation (XGPU_ERR_UNSUPPORTED). */
enum { XGPU_OP_DT = 3 };
int xgpu_elmat_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* elems,
                  const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                  const double* tvals, const double* wts, const xgpu_term* terms_host, int nterms,
                  int accumulate, const int32_t* slots, double* mats);
/* space-time element vectors (SymbolicCutLinearFormIntegrator with time_order >= 0, xfem/symboliccutlfi.cpp:51-140):
 * f(x, tref) * op_test(v), f = coef * h^hpow * x^e0 y^e1 z^e2 tref^e3, op_test in {ID, DT}.  vecs: n x nd. */
typedef struct { int op_test; double coef; int hpow; int expo[4]; } xgpu_stlterm;
int xgpu_elvec_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* elems,
                  const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                  const double* tvals, const double* wts, const xgpu_stlterm* terms_host, int nterms,
                  int accumulate, const int32_t* slots, double* vecs);

/* ---------------------------------------------------------------- (d) ghost penalty
 * SymbolicFacetPatchBilinearFormIntegrator::T_CalcFacetMatrix (xfem/symboliccutbfi.cpp:
 * 1982-2377) for coef*h^hpow*(u-u.Other())*(v-v.Other()); order<0 -> 2*fe_order (:2004-2012).
 * mats: n x (2nd) x (2nd), layout [el1 dofs ; el2 dofs] (:2207-2230). */
int xgpu_facetpatch(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* facets,
                    const int32_t* facet_els, const int32_t* el_vtx, const double* vtx_xyz,
                    int order, double coef, int hpow, double* mats);

/* dFacetPatch(downscale = s): SymbolicFacetPatchBilinearFormIntegrator::SetIRScaling (xfem/python_xfem.cpp:966-967).  State
 * of the context read by xgpu_facetpatch / xgpu_facetpatch_st: the volume rule of each element of the patch is scaled by s
 * towards the midpoint of the common facet and its weights by s^2 (ScaleAndMapIntegrationRule, xfem/symboliccutbfi.cpp:
 * 1849-1885, applied at :2019-2029).  XGPU_ERR_ARG unless 0 < s <= 1 (:2031-2033).  Default 1. */
int xgpu_facetpatch_set_downscale(xgpu_ctx* ctx, double s);

/* space-time ghost penalty: SymbolicFacetPatchBilinearFormIntegrator with time_order >= 0 (tensor-product rule in time,
 * xfem/symboliccutbfi.cpp:2095-2160) or a fixed reference time (time_order < 0: tref, :2162-2177) on a SpaceTimeFE.  On an
 * undeformed mesh the patch map does not depend on time (:2038-2040), so the matrix is (time matrix) kron (spatial patch
 * matrix of xgpu_facetpatch).  mats: n x (2 nd) x (2 nd), nd = nd_s*(time_fe_order+1), layout [el1 dofs ; el2 dofs], each
 * time-major. */
int xgpu_facetpatch_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* facets,
                       const int32_t* facet_els, const int32_t* el_vtx, const double* vtx_xyz, int order,
                       int time_order, double tref, double coef, int hpow, double* mats);

/* ---------------------------------------------------------------- (e) sparse assembly
 * RestrictedBilinearForm::GetGraph (utils/restrictedblf.cpp:142-163) + NGSolve MatrixGraph /
 * SparseMatrix::AddElementMatrix.  A batch is a table of dof rows (nrows x width, -1 = unused)
 * with one width x width matrix per row.  The pattern is the union of the row cliques of all
 * batches, columns ascending, explicit zeros kept; values are summed in a fixed order
 * (batch, then row index) so results are bit-reproducible run to run. */
typedef struct {
  const int32_t* table;
  const double* mats;
  int64_t nrows;
  int width;
  const double* coefs; /* xgpu_csr_assemble_fused only: procedural affine batch (xgpu_affine_coefs), mats NULL; else NULL */
} xgpu_batch;
int xgpu_csr_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                  xgpu_plan** plan_out, int64_t* nnz_host);
int xgpu_csr_pattern(xgpu_ctx* ctx, xgpu_plan* plan, int64_t* rowptr, int32_t* colidx);
int xgpu_csr_scatter(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches,
                     const int64_t* rowptr, const int32_t* colidx, double* vals);
/* the transposed table alone (no sparsity pattern): enough for xgpu_vec_scatter */
int xgpu_adjacency_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                        xgpu_plan** plan_out);
/* LinearForm assembly (NGSolve AddElementVector): out[d] = sum of the element-vector entries whose dof is d, in
 * (batch, row, local index) order — deterministic.  A batch's `mats` points to nrows x width doubles here. */
int xgpu_vec_scatter(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches, double* out);
/* the same with the MEAN over the contributing entries instead of the sum (dofs without entries: 0): out[d * stride + comp].
 * NGSolve's GridFunction.Set and ProjectShift average element-local projections this way (lsetcurving/projshift.cpp:165-197) */
int xgpu_vec_scatter_mean(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches, double* out,
                          int stride, int comp);
/* xgpu_csr_pattern + xgpu_csr_scatter in one pass over the rows (BilinearForm.Assemble builds both every time):
 * the scatter kernels read each row's columns from the plan's pattern scratch and write rowptr, colidx AND vals
 * — the separate compaction copy (and its 1.6 GB of DRAM traffic on config 2) disappears.  Same results, bit for bit. */
int xgpu_csr_assemble(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches,
                      int64_t* rowptr, int32_t* colidx, double* vals);
/* The one-row-pass path.  xgpu_csr_plan_rows = transposed table + a lean COUNT pass (distinct columns per row, keys
 * only) + scan: the plan knows rowptr and nnz.  xgpu_csr_assemble_fused then finishes every row in ONE visit: its
 * contributions are deduplicated and summed in (batch, row, local index) order in a shared-memory table sized from the
 * row's exact length, sorted, and written to their final CSR position (rowptr, colidx, vals; colidx / vals hold nnz
 * entries).  A batch with `coefs` set is a PROCEDURAL batch of uncut affine elements (SymbolicCutBFI on elements where
 * the cut rule is the plain rule, xfem/symboliccutbfi.cpp:127-135): row i of its element matrix is
 * sum_m coefs[e][m] * tref[m][i][:]  with the reference tensors of xgpu_reftensor — bit-identical to xgpu_elmat_std's
 * affine path, without the matrices ever reaching memory.  XGPU_ERR_UNSUPPORTED from xgpu_csr_plan_rows (duplicate dofs
 * in a table row, rows wider than 64, rows with too many candidates): use xgpu_csr_plan + xgpu_csr_assemble.  Results
 * equal the two-pass path bit for bit (rows up to 32 wide; wider rows: same pattern, strict sequential summation order). */
int xgpu_csr_plan_rows(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                       xgpu_plan** plan_out, int64_t* nnz_host);
/* the count pass alone, on an existing transposed table (xgpu_adjacency_plan, xgpu_restrict_plan) */
int xgpu_csr_count_rows(xgpu_ctx* ctx, xgpu_plan* plan, int64_t* nnz_host);
/* Restrict + dof rows + transposed table in one sequence (RestrictedFESpace utils/restrictedfespace.cpp:23-73 followed by
 * the table of RestrictedBilinearForm::GetGraph, utils/restrictedblf.cpp:142-163): the active dofs are the dofs of the
 * listed elements (nlists element lists, lists_host[i] = DEVICE pointer to counts_host[i] element numbers, NULL = 0..count-1);
 * dofmap[d] (ndof_full entries) = compressed number of dof d or -1, *ndof_host = number of active dofs,
 * tables_host[i] (DEVICE pointers, counts_host[i] x nd int32) = the compressed dof rows of list i = batch i of the plan.
 * Equivalent to xgpu_dofs_of_elements + xgpu_bitset_rank + xgpu_gather_rows + xgpu_adjacency_plan with one atomic pass
 * and one table pass less. */
int xgpu_restrict_plan(xgpu_ctx* ctx, int nd, const int32_t* el_dofs, int nlists, const int32_t* const* lists_host,
                       const int64_t* counts_host, int64_t ndof_full, int32_t* dofmap, int64_t* ndof_host,
                       int32_t* const* tables_host, xgpu_plan** plan_out);
int xgpu_csr_assemble_fused(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches,
                            const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals);
/* the same for the rows [row_lo, row_hi) only (colidx / vals of the other rows are not touched; copy_rowptr != 0 also copies
 * the row offsets).  An element-partitioned assembly finishes the rows it has to SEND first and lets them travel while the
 * rows it owns are summed.  XGPU_ERR_UNSUPPORTED when the plan has rows outside the thread-per-row / warp-per-row kernels. */
int xgpu_csr_assemble_fused_rows(xgpu_ctx* ctx, xgpu_plan* plan, const xgpu_batch* batches_host, int nbatches,
                                 const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals,
                                 int64_t row_lo, int64_t row_hi, int copy_rowptr);
void xgpu_plan_destroy(xgpu_plan* plan);
/* device time of the two halves of xgpu_csr_plan (CUDA events on the plan's stream): ms_host[0] =
 * transposed table (histogram, scan, fill, segment sort), ms_host[1] = pattern pass (bench.py) */
int xgpu_plan_phase_ms(xgpu_plan* plan, double* ms_host);
/* ---------------------------------------------------------------- (e') multi-GPU interface rows
 * Element-partitioned assembly: rows whose dof is shared with a neighbouring rank are summed at the owner.  Both sides
 * hold identical column orders for such rows, so the message is the packed VALUES of the listed rows (rows[i] < 0: an
 * inactive candidate, length 0).  The send/recv itself (NCCL over NVLink) is issued by the host layer on `out` / `in`.
 * offsets: n+1 int64 (device) = exclusive scan of the row lengths; total_host may be NULL (no synchronisation). */
int xgpu_rows_offsets(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, int64_t* offsets,
                      int64_t* total_host);
int xgpu_rows_pack(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, const int64_t* offsets,
                   const double* vals, double* out);
int xgpu_rows_add(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, const int64_t* offsets,
                  const double* in, double* vals);
/* ---------------------------------------------------------------- (f4) cut integrals over facets / element boundaries
 * SymbolicCutFacetBilinearFormIntegrator::T_CalcFacetMatrix (xfem/symboliccutbfi.cpp:886-1290, dCut(..., skeleton=True)) and
 * SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixEBAdd (:488-883, dCut(..., element_boundary=True)) on affine simplices.
 * The rule of every listed facet comes from xgpu_cutquad_* applied to the FACET-vertex table (dimension dim-1: points in the
 * facet's reference coordinates, reference weights; offsets n+1).  facet_vtx: n x dim global vertex numbers in the order the
 * rule was built for; facet_els: n x 2 (nsides = 1: column 1 unused).  Integrand menu: sum of
 * coef * h^hpow * Op_test(v|side) * Op_trial(u|side), kind 0 = value, 1 = nF . grad with nF the outward unit normal of the
 * FIRST element, side 0 = that element, 1 = .Other(); h = mesh size of the first element.  mats: n x (nsides nd)^2 row-major,
 * rows = test, dofs of side 0 first — written, not accumulated. */
typedef struct { int trial_kind, trial_side, test_kind, test_side; double coef; int hpow; } xgpu_fterm;
int xgpu_facet_elmat(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* facet_vtx, const int32_t* facet_els,
                     const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                     const double* wts, const xgpu_fterm* terms_host, int nterms, int nsides, double* mats);
/* The interface bookkeeping of a step over the STATIC candidate tables of a partition (ngsxfem_b200/distributed.py), without
 * host round trips.  xgpu_iface_marks: flags[pair_gpos[i]] = 1 for every static (element, flag position) pair whose element
 * is active (bit of has_bits = HASNEG BitArray words); flags (nflags bytes) is cleared first — the reference's
 * AllReduceNodalData(..., MPI_LOR) over shared dofs (xfem/cutinfo.cpp:194-199) restricted to the candidates.
 * xgpu_iface_rows: rows[i] = flags[i] ? dofmap[cand[i]] : -1 (the row lists xgpu_rows_* take).  xgpu_iface_index: the
 * owner-local index message idx[i] = flags[i] ? dofmap[cand[i]] - n_lo : -1.  xgpu_l2g_own / xgpu_l2g_ghosts: global dof
 * numbers = (sum of counts[k], k < owner) + owner-local index; counts = the all-reduced owned counts (device). */
int xgpu_iface_marks(xgpu_ctx* ctx, const uint32_t* has_bits, int64_t npairs, const int32_t* pair_elem,
                     const int32_t* pair_gpos, int64_t nflags, uint8_t* flags);
int xgpu_iface_rows(xgpu_ctx* ctx, int64_t n, const int32_t* cand, const uint8_t* flags, const int32_t* dofmap,
                    int32_t* rows);
int xgpu_iface_index(xgpu_ctx* ctx, int64_t n, const int32_t* cand, const uint8_t* flags, const int32_t* dofmap,
                     int64_t n_lo, int64_t* idx);
int xgpu_l2g_own(xgpu_ctx* ctx, int64_t ndof, int64_t n_lo, int64_t n_own, const int64_t* counts, int rank, int64_t* l2g);
int xgpu_l2g_ghosts(xgpu_ctx* ctx, int64_t n, const int32_t* dofmap_range, const int64_t* rbuf, const int64_t* counts,
                    int q, int64_t* l2g);
/* y = A x for checks (CSR SpMV, one warp per row) */
int xgpu_csr_spmv(xgpu_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* colidx,
                  const double* vals, const double* x, double* y);

/* ---------------------------------------------------------------- isoparametric deformation producer (row f3)
 * LevelSetMeshAdaptation.CalcDeformation (python/lsetcurv.py:281-356): lset_ho.Set(levelset); qn.Set(grad lset_ho);
 * InterpolateToP1; ProjectShift (lsetcurving/projshift.cpp:9-201).  H1 functions are coefficient vectors of the hierarchical
 * basis (el_dofs: ne x xgpu_h1_ndof(dim, fe_order), all entries >= 0), affine mesh (the deformation is what is being made). */
/* ElementInRelevantBand for every element (lsetcurving/calcpointshift.cpp:120-141) -> BitArray words */
int xgpu_band_elements(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* lset_p1,
                       double lower, double upper, uint32_t* bits_out);
/* physical coordinates of the points of SelectIntegrationRule(order_rule) on the listed elements: out n x np x dim
 * (np = xgpu_rule_size) — where a user level-set function is sampled for GridFunction.Set */
int xgpu_rule_points_phys(xgpu_ctx* ctx, int dim, int order_rule, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                          const double* vtx_xyz, double* out);
/* physical gradient of the H1(fe_order) function `vals` at those points: out n x np x dim (lset_ho.Deriv()) */
int xgpu_h1_grad_points(xgpu_ctx* ctx, int dim, int fe_order, int order_rule, int64_t n, const int32_t* elems,
                        const int32_t* el_vtx, const double* vtx_xyz, const int32_t* el_dofs, const double* vals, double* out);
/* element-local L2 projection of point values (fvals n x np x ncomp at the points of order_rule >= 2 fe_order) onto
 * H1(fe_order): elres ncomp x n x nd, component-major — one batch per component for xgpu_vec_scatter_mean.  This is the
 * element half of GridFunction.Set and of ProjectShift (projshift.cpp:92-110; zero_vertex != 0 zeroes the vertex rows, :108) */
int xgpu_l2_project(xgpu_ctx* ctx, int dim, int fe_order, int order_rule, int64_t n, const int32_t* elems,
                    const int32_t* el_vtx, int ncomp, const double* fvals, int zero_vertex, double* elres);
/* ShiftIntegrator::CalcElementVector, pointwise part (lsetcurving/shiftintegrators.cpp:69-121, SearchCorrespondingPoint
 * calcpointshift.cpp:145-206): for every point of the rule of order 2 fe_order on the listed elements the physical shift
 * F (xhat* - xhat), xhat* on the line xhat + s F^{-1} qn with lset_ho(xhat*) = (1-alpha) lset_p1(xhat) + alpha lset_ho(xhat),
 * |xhat* - xhat| clamped to `threshold` (< 0: none).  qn_mode 0: grad lset_p1 (qn == nullptr in the reference; exact gradient
 * instead of its 1e-7 central difference), 1: grad lset_ho at the point (an L2(order)^dim qn set from lset_ho.Deriv()),
 * 2: `qn` = coefficients ndof x dim of an H1(fe_order)^dim field.  blending (nullable): alpha at the points, n x np.
 * shift_out n x np x dim; *nfail_host (nullable, synchronises) = searches that did not converge in 20 steps (shift 0). */
int xgpu_shift_points(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                      const double* vtx_xyz, const int32_t* el_dofs, const double* lset_ho, const double* lset_p1,
                      int qn_mode, const double* qn, const double* blending, double threshold, double* shift_out,
                      int64_t* nfail_host);

/* ---------------------------------------------------------------- general coefficient functions
 * NGSolve evaluates a CoefficientFunction tree at every quadrature point (xfem/symboliccutbfi.cpp:174-193,
 * xfem/symboliccutlfi.cpp:96-118).  Here a coefficient is TABULATED at the points of the per-element rules and folded into
 * the rule weights (B^T diag(w D) B and sum_q w f phi_i only see the product), so every kernel above takes it unchanged:
 * xgpu_rule_phys gives the physical coordinates (and measures |det F(xhat_q)|) of the points — deformation-aware —, the
 * caller evaluates its function there, multiplies the weights, and calls xgpu_elmat_rule / xgpu_elvec_rule / the integrals.
 * xgpu_h1_eval_rule evaluates a GridFunction of H1(fe_order) at the same points (el_dofs entries < 0: coefficient 0). */
int xgpu_rule_phys(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                   const int64_t* offsets, const double* pts, double* xyz_out, double* meas_out);
int xgpu_h1_eval_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                      const int32_t* el_dofs, const double* vals, const int64_t* offsets, const double* pts, double* out);
/* fixed-order (tile tree) sum of n doubles: bit-reproducible; synchronises (Integrate with a tabulated coefficient) */
int xgpu_reduce_sum(xgpu_ctx* ctx, int64_t n, const double* vals, double* result_host);

#ifdef __cplusplus
}
#endif
#endif
