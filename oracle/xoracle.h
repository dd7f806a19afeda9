/* xoracle.h — CPU restatement ORACLE of ngsxfem's cut-element assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ngsxfem_b200/ (the product) may include, link,
 * import or execute anything under oracle/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker / baseline.
 *
 * Every function cites the reference file:line (relative to the ngsxfem source tree) whose
 * algorithm it restates.  The reference itself cannot be compiled here: every .cpp includes
 * NGSolve headers (<comp.hpp>, <fem.hpp>) and NGSolve is not installed.
 *
 * PARITY PIN STATUS
 *   pinned   : the cut geometry / quadrature construction / classification logic — checked
 *              against every analytic known-answer value the reference's own tests hold for
 *              this path (tests/test_oracle_kat.py lists them with reference test file:line).
 *   UNPINNED : the arithmetic living in NGSolve (third party, pip `ngsolve>=6.2.2606`, not
 *              vendored): reference quadrature tables (SelectIntegrationRule), the H1HighOrderFE
 *              shape functions, dof / edge / face numbering.  Those are restated from NGSolve's
 *              published algorithm and are injectable data tables on the product side.
 */
#ifndef XORACLE_H
#define XORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* utils/ngsxstd.hpp:16-30 */
enum { XO_NEG = 0, XO_POS = 1, XO_IF = 2 };
enum { XO_CDOM_NO = 0, XO_CDOM_NEG = 1, XO_CDOM_POS = 2, XO_CDOM_UNCUT = 3, XO_CDOM_IF = 4,
       XO_CDOM_HASNEG = 5, XO_CDOM_HASPOS = 6, XO_CDOM_ANY = 7 };

/* error codes (the reference throws ngcore::Exception at these places) */
enum {
  XO_OK = 0,
  XO_ERR_PLAIN_GEOM = -1,   /* straightcutrule.cpp:45-49  "cut a simplex with a plain geometry lset" */
  XO_ERR_BAD_SCUT = -2,     /* straightcutrule.cpp:144-149 "Bad length of s_cut!" */
  XO_ERR_NOT_IMPL = -3,     /* straightcutrule.cpp:190-194 "Cutting this part ... not implemented" */
  XO_ERR_CAPACITY = -4,
  XO_ERR_ARG = -5
};

#define XO_MAXPTS_RULE 2048

/* ------------------------------------------------------------------ rules (NGSolve restated) */
/* number of points of SelectIntegrationRule(ET of dimension dim, order) */
int xo_rule_size(int dim, int order);
/* points: n x 3 (unused coords 0), weights: n.  returns n */
int xo_select_rule(int dim, int order, double* pts, double* wts);
/* Gauss-Jacobi on [0,1] with weight (1-x)^alpha, n points (ascending x) */
void xo_gauss_jacobi01(int n, int alpha, double* x, double* w);

/* ------------------------------------------------------------------ cut geometry */
typedef struct { int D; double p[4][3]; } xo_simplex;

int xo_check_if_straight_cut(const double* v, int n, double eps);
void xo_ref_simplex(int dim, xo_simplex* s);
double xo_simplex_volume(const xo_simplex* s);
/* LevelsetCutSimplex::Decompose; out must hold 3 simplices; returns count or error (<0) */
int xo_decompose(const xo_simplex* s, const double* lsetvals, int dt, xo_simplex* out);

/* element transformation: affine simplex, optionally + H1 deformation of order `def_order`
 * (def: ndof x dim coefficients in the oracle's own hierarchical basis, xo_h1_shapes) */
typedef struct {
  int dim;
  double vtx[4][3];
  int vnum[4];
  int def_order;
  const double* def;
} xo_trafo;
/* F is dim x dim row-major: F[i][j] = d x_i / d xhat_j ; x = mapped point */
void xo_trafo_eval(const xo_trafo* T, const double* xhat, double* F, double* x);

/* StraightCutIntegrationRule (straightcutrule.cpp:582-677).
 * returns: number of points (>0), 0 if the reference returns nullptr, <0 on error.
 * *is_std is set to 1 when the (shared) standard rule of the uncut element is returned. */
int xo_straight_cut_rule(const xo_trafo* T, const double* lsetvals, int dt, int order,
                         double* pts, double* wts, int cap, int* is_std);
/* multi level set: lsets is (dim+1) x M row-major (column j = level set j);
 * dts: ntuples x M.  (mlsetintegration.cpp:6-186, straightcutrule.cpp:679-812) */
int xo_mlset_cut_rule(const xo_trafo* T, const double* lsets, int M, const int* dts, int ntuples,
                      int order, double* pts, double* wts, int cap);
/* space-time (spacetimecutrule.cpp:8-244): lset_st is time-major (nt x (dim+1)), time nodal
 * basis of order time_fe_order on [0,1]; output points carry t in tvals. */
int xo_spacetime_cut_rule(const xo_trafo* T, const double* lset_st, int time_fe_order, int dt,
                          int order_time, int order_space, double* pts, double* tvals,
                          double* wts, int cap);
void xo_time_shapes(int order, double t, double* shape);

/* ------------------------------------------------------------------ CutInfo (xfem/cutinfo.cpp) */
/* CutInformation::Update for simplices of dimension dim (cutinfo.cpp:61-133).
 * bits_*: packed bytes, bit i = byte i>>3, bit i&7 (NGSolve BitArray layout).
 * returns 0 or the first error code; *err_elem = first failing element. */
int xo_cutinfo_update(int dim, int64_t ne, const int32_t* el_vtx, const double* phi,
                      uint8_t* bits_neg, uint8_t* bits_pos, uint8_t* bits_if,
                      double* cut_ratio, int64_t* err_elem);
/* MultiLevelsetCutInformation::UpdateElementsOfDomainType (cutinfo.cpp:356-398) */
void xo_ml_elements_of_type(int dim, int64_t ne, const int32_t* el_vtx, const double* phis /*M x nv*/,
                            int64_t nv, int M, const int* dts, int ntuples, uint8_t* bits);
/* ...UpdateElementsWithContribution (cutinfo.cpp:306-334); order as the reference: -1 */
int xo_ml_elements_with_contribution(int dim, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz,
                                     const double* phis, int64_t nv, int M, const int* dts,
                                     int ntuples, uint8_t* bits);
/* GetFacetsWithNeighborTypes (cutinfo.cpp:476-538), no periodic identification.
 * facet_els: nf x 2, -1 for no neighbour. */
void xo_facets_with_neighbor_types(int64_t nf, const int32_t* facet_els, const uint8_t* a,
                                   const uint8_t* b, int bnd_val_a, int bnd_val_b, int use_and,
                                   uint8_t* out);
void xo_facets_with_neighbor_types_periodic(int64_t nf, const int32_t* facet_els, const int32_t* periodic_facet,
                                            const uint8_t* a, const uint8_t* b, int bnd_val_a, int bnd_val_b, int use_and,
                                            uint8_t* out);
/* GetElementsWithNeighborFacets (cutinfo.cpp:540-566) */
void xo_elements_with_neighbor_facets(int64_t nf, const int32_t* facet_els, const uint8_t* facets,
                                      int64_t ne, uint8_t* out);
/* GetDofsOfElements (cutinfo.cpp:599-622) */
void xo_dofs_of_elements(int64_t ne, int nd, const int32_t* el_dofs, const uint8_t* elems,
                         int64_t ndof, uint8_t* out);

/* ------------------------------------------------------------------ H1 basis (NGSolve restated, unpinned) */
int xo_h1_ndof(int dim, int order);
/* shape: nd, dshape: nd x dim (reference gradients) at xhat; vnum = global vertex numbers */
void xo_h1_shapes(int dim, int order, const int* vnum, const double* xhat, double* shape, double* dshape);

/* ------------------------------------------------------------------ forms */
/* one term:  coef * h^hpow * (op_test phi_i) * (op_trial phi_j) ;  GRAD x GRAD = dot product,
 * DN = normal derivative with n = grad(lset_p1)/|grad(lset_p1)| (physical). */
enum { XO_OP_ID = 0, XO_OP_GRAD = 1, XO_OP_DN = 2 };
typedef struct { int op_trial, op_test; double coef; int hpow; } xo_term;

/* SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixAdd (symboliccutbfi.cpp:76-486).
 * elmat (nd x nd row-major, rows = test, cols = trial) is ADDED into.  rule order `order`
 * (<0: 2*fe_order minus the minimal diff orders, :100-118).  returns #points used or error. */
int xo_cut_elmat(const xo_trafo* T, int fe_order, const double* lsetvals, int dt, int order,
                 const xo_term* terms, int nterms, double* elmat);
/* the same with an explicit rule (points in reference coords, weights w/o |det F|) */
void xo_elmat_from_rule(const xo_trafo* T, int fe_order, const double* lsetvals,
                        const double* pts, const double* wts, int np,
                        const xo_term* terms, int nterms, double* elmat);
/* cut integrals over interior facets / element boundaries (xo_facet.c; symboliccutbfi.cpp:886-1290, :488-883).
 * kind: 0 = value, 1 = nF . grad (nF = outward unit normal of the first element); side: 0 = this element, 1 = .Other() */
typedef struct { int trial_kind, trial_side, test_kind, test_side; double coef; int hpow; } xo_fterm;
int xo_cut_facet_elmat(const xo_trafo* T1, const xo_trafo* T2, int fe_order, const int* facet_vnum, const double* lset_facet,
                       int dt, int order, const xo_fterm* terms, int nterms, double* mat);
/* space-time FE (spacetime/SpaceTimeFE.cpp:57-135, time-major tensor basis) on a space-time rule: ops ID, GRAD.GRAD,
 * XO_OP_DT (reference-time derivative).  nd = nd_s*(time_fe_order+1); ADDS into elmat / elvec. */
enum { XO_OP_DT = 3 };
void xo_time_dshapes(int order, double t, double* dshape);
void xo_elmat_st_from_rule(const xo_trafo* T, int fe_order, int time_fe_order, const double* pts, const double* tvals,
                           const double* wts, int np, const xo_term* terms, int nterms, double* elmat);
typedef struct { int op_test; double coef; int hpow; int expo[4]; } xo_stlterm;
void xo_elvec_st_from_rule(const xo_trafo* T, int fe_order, int time_fe_order, const double* pts, const double* tvals,
                           const double* wts, int np, const xo_stlterm* terms, int nterms, double* elvec);
/* SymbolicFacetPatchBilinearFormIntegrator::T_CalcFacetMatrix (symboliccutbfi.cpp:1982-2377)
 * form: coef * h^hpow * (u-u.Other())*(v-v.Other()); elmat (2nd x 2nd) is ZEROED then filled.
 * returns number of Newton failures. */
int xo_facet_patch_elmat(const xo_trafo* T1, const xo_trafo* T2, int fe_order, int order,
                         double coef, int hpow, double* elmat);
/* with dFacetPatch(downscale = s), ScaleAndMapIntegrationRule (symboliccutbfi.cpp:1849-1885, :2019-2033); -1 unless 0 < s <= 1 */
int xo_facet_patch_elmat_scaled(const xo_trafo* T1, const xo_trafo* T2, int fe_order, int order,
                                double coef, int hpow, double downscale, double* elmat);

/* SymbolicCutLinearFormIntegrator::T_CalcElementVector (xfem/symboliccutlfi.cpp:51-140) for the closed menu
 *   f(x) * op_test(v),  f = coef * h^hpow * x^e0 y^e1 z^e2,  op_test in {ID, DN}:
 * elvec_i += sum_q (w_q |det F|) f(x_q) op_test(phi_i)(x_q).  Rule order: `order`, or 2*fe_order if < 0
 * (:86-88; no reduction by the differentiation order).  returns #points used or an error. */
typedef struct { int op_test; double coef; int hpow; int expo[3]; } xo_lterm;
void xo_elvec_from_rule(const xo_trafo* T, int fe_order, const double* lsetvals,
                        const double* pts, const double* wts, int np,
                        const xo_lterm* terms, int nterms, double* elvec);
int xo_cut_elvec(const xo_trafo* T, int fe_order, const double* lsetvals, int dt, int order,
                 const xo_lterm* terms, int nterms, double* elvec);
int xo_batch_cut_elvecs(int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                        const double* vtx_xyz, const double* phi, int dt, int order,
                        const xo_lterm* terms, int nterms, int def_order, const double* def,
                        double* vecs);

/* ------------------------------------------------------------------ sparse assembly */
/* Graph = union of cliques of the rows of `table` (nrows x width, -1 = unused), sorted columns
 * (restrictedblf.cpp:142-163 -> NGSolve MatrixGraph).  Call with colidx==NULL to get nnz via rowptr. */
int64_t xo_graph(int64_t ndof, int64_t nrows, int width, const int32_t* table,
                 int64_t* rowptr, int32_t* colidx);
/* timing variant of xo_graph (one sort per row, one call): *colidx_out is malloc'ed, release with xo_free */
int64_t xo_graph_once(int64_t ndof, int64_t nrows, int width, const int32_t* table, int64_t* rowptr, int32_t** colidx_out);
void xo_free(void* p);
/* vals += scatter of `mats` (nrows x width x width) by table rows (AddElementMatrix) */
void xo_scatter(int64_t ndof, int64_t nrows, int width, const int32_t* table, const double* mats,
                const int64_t* rowptr, const int32_t* colidx, double* vals);
/* timing-only variant: table rows spread over the host threads, atomic adds (order not fixed) */
void xo_scatter_threads(int64_t ndof, int64_t nrows, int width, const int32_t* table, const double* mats,
                        const int64_t* rowptr, const int32_t* colidx, double* vals);

#ifdef __cplusplus
}
#endif
#endif

/* ------------------------------------------------------------------ isoparametric deformation producer (xo_shift.c)
 * lsetcurving/projshift.cpp:9-201, shiftintegrators.cpp:39-124, calcpointshift.cpp:81-206; python/lsetcurv.py:281-356 */
int xo_element_in_band(const double* vals, int n, double lower, double upper);
int xo_rule_points_phys(int dim, int order_rule, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, double* out);
void xo_set_from_points(int dim, int fe_order, int order_rule, int64_t ne, const int32_t* el_vtx, const int32_t* el_dofs,
                        int64_t ndof, int ncomp, const double* fvals, double* out);
void xo_set_gradient(int dim, int fe_order, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, const int32_t* el_dofs,
                     int64_t ndof, const double* lset_ho, double* qn_out);
int64_t xo_project_shift(int dim, int fe_order, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz,
                         const int32_t* el_dofs, int64_t ndof, const double* lset_ho, const double* lset_p1, int qn_mode,
                         const double* qn, const uint8_t* ba, const double* blending, double lower, double upper,
                         double threshold, double* deform);
double xo_integrate_one_deformed(int dim, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int dt,
                                 int order, int def_order, const int32_t* el_dofs_def, const double* def_vals, int* err);
