// integration/gpurestrictedblf.cpp — the NGSolve-side adapter of RestrictedBilinearForm::Assemble for the
// BASELINE config-2 shape: one P1 level set, H1 order <= 2,  (grad u . grad v + u v) dCut(NEG) on HASNEG  +  u v dCut(IF).
// NGSolve drives assembly by one virtual call per element (xfem/symboliccutbfi.cpp:48-58), which cannot feed a GPU; the
// adapter overrides DoAssemble, runs the batched ABI sequence of ngsxfem_b200/pipeline.py and writes the values into NGSolve's
// own SparseMatrix, whose graph (utils/restrictedblf.cpp:142-163) has the same rows and ascending columns.
#include <comp.hpp>
#ifdef XGPU_ADAPTER_STUBS
#include "xfem_stubs.hpp"
#else
#include "restrictedblf.hpp"
#endif
#include "gpu_common.hpp"

namespace ngcomp {
using xgpu_adapter::Check;
using xgpu_adapter::DeviceBuffer;

class GpuRestrictedBilinearForm : public RestrictedBilinearForm<double, double> {
  xgpu_ctx* ctx = nullptr;
  shared_ptr<GridFunction> lsetp1;
  DeviceBuffer<int32_t> d_elvtx, d_eldofs, d_lneg, d_lif, d_tneg, d_tif, d_dofmap, d_col;
  DeviceBuffer<double> d_xyz, d_phi, d_ratio, d_coefs, d_mif, d_vals;
  DeviceBuffer<uint32_t> d_neg, d_pos, d_if;
  DeviceBuffer<int64_t> d_rowptr;
public:
  GpuRestrictedBilinearForm(shared_ptr<GridFunction> alset) : lsetp1(alset)
  {
    if (xgpu_create(0, &ctx) != XGPU_OK) throw Exception("xgpu: no CUDA device (there is no CPU fallback)");
  }
  ~GpuRestrictedBilinearForm() override { if (ctx) xgpu_destroy(ctx); }

  void DoAssemble(LocalHeap& lh) override
  {
    auto ma = fespace->GetMeshAccess();
    const int dim = ma->GetDimension(), p = fespace->GetOrder();
    if (dim != 3 || p < 1 || p > 2) throw Exception("GpuRestrictedBilinearForm: tets, H1 order 1 or 2 (other shapes: xgpu_elmat_* + xgpu_csr_plan)");
    const int64_t ne = (int64_t)ma->GetNE(VOL), nv = (int64_t)ma->GetNV(), nd = xgpu_h1_ndof(dim, p);
    // (1) mesh / space tables -> device (once per mesh in a real adapter; every call here for brevity)
    Array<int32_t> elvtx(4 * ne), eldofs(nd * ne);
    Array<double> xyz(3 * nv);
    Array<int> dn;
    for (size_t e : Range(ne)) {
      auto v = ma->GetElement(ElementId(VOL, e)).Vertices();
      for (int k = 0; k < 4; k++) elvtx[4 * e + k] = v[k];
      fespace->GetDofNrs(ElementId(VOL, e), dn);          // NGSolve's numbering (vertices, edges): -1 = unused dof
      for (int k = 0; k < nd; k++) eldofs[nd * e + k] = k < (int)dn.Size() ? dn[k] : -1;
    }
    for (size_t i : Range(nv)) { Vec<3> pt; ma->GetPoint<3>(i, pt); for (int c = 0; c < 3; c++) xyz[3 * i + c] = pt(c); }
    d_elvtx.Upload(elvtx); d_eldofs.Upload(eldofs); d_xyz.Upload(xyz);
    auto phi = lsetp1->GetVector().FVDouble();
    d_phi.Upload(phi.Data(), phi.Size());
    // (2) CutInfo + the two element lists
    const size_t nw = (ne + 31) / 32;
    d_neg.Resize(nw); d_pos.Resize(nw); d_if.Resize(nw); d_ratio.Resize(ne);
    Check(ctx, xgpu_classify(ctx, dim, ne, d_elvtx, d_phi, d_neg, d_pos, d_if, d_ratio));
    xgpu_adapter::CheckGeometry(ctx);
    int64_t cnt[2] = {0, 0};
    const uint32_t* sets[2] = {d_neg, d_if};
    const int64_t lens[2] = {ne, ne};
    Check(ctx, xgpu_bitset_count_many(ctx, 2, sets, lens, cnt));    // one host wait for both list lengths
    const int64_t n_neg = cnt[0], n_if = cnt[1];
    d_lneg.Resize(n_neg); d_lif.Resize(n_if);
    Check(ctx, xgpu_bitset_to_list(ctx, d_neg, ne, d_lneg));
    Check(ctx, xgpu_bitset_to_list(ctx, d_if, ne, d_lif));
    // (3) Restrict + dof rows + transposed table in one sequence; the compressed numbering equals NGSolve's
    //     CompressedFESpace numbering (ascending over the active dofs, utils/restrictedfespace.cpp:23-51)
    const int64_t ndof_full = (int64_t)fespace->GetNDof();
    d_dofmap.Resize(ndof_full); d_tneg.Resize(n_neg * nd); d_tif.Resize(n_if * nd);
    const int32_t* lists[2] = {d_lneg, d_lif};
    int32_t* tables[2] = {d_tneg, d_tif};
    const int64_t counts[2] = {n_neg, n_if};
    int64_t ndof = 0, nnz = 0;
    xgpu_plan* plan = nullptr;
    Check(ctx, xgpu_restrict_plan(ctx, (int)nd, d_eldofs, 2, lists, counts, ndof_full, d_dofmap, &ndof, tables, &plan));
    Check(ctx, xgpu_csr_count_rows(ctx, plan, &nnz));
    // (4) cut element matrices, the cut rule made inside the kernel as T_CalcElementMatrixAdd does (NEG volume terms and the IF
    //     term in one pass; orders > 2 / deformed meshes: xgpu_cutquad_count/_fill + xgpu_elmat_rule_pair), 7 coefficients per
    //     uncut element
    const xgpu_term vol[2] = {{XGPU_OP_GRAD, XGPU_OP_GRAD, 1.0, 0}, {XGPU_OP_ID, XGPU_OP_ID, 1.0, 0}}, ifc[1] = {{XGPU_OP_ID, XGPU_OP_ID, 1.0, 0}};
    const int order = 2 * p;
    d_mif.Resize(n_if * nd * nd); d_coefs.Resize(n_neg * 8);
    Check(ctx, xgpu_elmat_cut_pair(ctx, dim, p, n_if, d_lif, d_elvtx, d_xyz, d_phi, XGPU_NEG, order, vol, 2, XGPU_IF, order, ifc, 1, 0, nullptr, d_mif));
    xgpu_adapter::CheckGeometry(ctx);
    Check(ctx, xgpu_affine_coefs(ctx, dim, n_neg, d_lneg, d_elvtx, d_xyz, vol, 2, d_coefs));
    const double* tref = nullptr; int ncoef = 0;
    Check(ctx, xgpu_reftensor(ctx, dim, p, order, &tref, &ncoef));
    // (5) one row pass: pattern + values
    xgpu_batch batches[2] = {{d_tneg, nullptr, n_neg, (int)nd, d_coefs}, {d_tif, d_mif, n_if, (int)nd, nullptr}};
    d_rowptr.Resize(ndof + 1); d_col.Resize(nnz); d_vals.Resize(nnz);
    Check(ctx, xgpu_csr_assemble_fused(ctx, plan, batches, 2, tref, ncoef, d_rowptr, d_col, d_vals));
    xgpu_plan_destroy(plan);
    // (6) into NGSolve's matrix: same rows (compressed dofs), same ascending columns -> the values copy straight over
    if ((int64_t)mat->Height() != ndof || (int64_t)mat->NZE() != nnz)
      throw Exception("GpuRestrictedBilinearForm: NGSolve's graph and the device graph differ (facet couplings? use xgpu_csr_plan with the facet tables)");
    d_vals.Download(mat->GetValues().Data(), nnz);
  }
};
}  // namespace ngcomp
