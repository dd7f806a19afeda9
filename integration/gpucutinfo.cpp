// integration/gpucutinfo.cpp — the NGSolve-side adapter of CutInfo: a CutInformation whose Update runs
// xgpu_classify and refills NGSolve's own BitArrays in place (xfem/cutinfo.cpp:61-261; pybind xfem/python_xfem.cpp:47-199).
// In ngsxfem this file includes the real <comp.hpp> and "cutinfo.hpp"; here it is syntax-checked against integration/stubs/.
#include <comp.hpp>
#ifdef XGPU_ADAPTER_STUBS
#include "xfem_stubs.hpp"
#else
#include "cutinfo.hpp"
#endif
#include "gpu_common.hpp"

namespace ngcomp {
using xgpu_adapter::DeviceBuffer;

class GpuCutInformation : public CutInformation {
  xgpu_ctx* ctx = nullptr;
  int dim = 0;
  size_t ne = 0;
  DeviceBuffer<int32_t> d_elvtx;        // static between level-set updates
  DeviceBuffer<double> d_phi, d_ratio;
  DeviceBuffer<uint32_t> d_neg, d_pos, d_if;
public:
  GpuCutInformation(shared_ptr<MeshAccess> ama) : CutInformation(ama)
  {
    if (xgpu_create(0, &ctx) != XGPU_OK) throw Exception("xgpu: no CUDA device (there is no CPU fallback)");
    dim = ma->GetDimension();
    ne = ma->GetNE(VOL);
    Array<int32_t> elvtx((dim + 1) * ne);
    for (size_t e : Range(ne)) {
      auto v = ma->GetElement(ElementId(VOL, e)).Vertices();
      if (v.Size() != size_t(dim + 1)) throw Exception("GpuCutInformation: simplicial meshes only (quads / hexes use the legacy path)");
      for (int k = 0; k < dim + 1; k++) elvtx[(dim + 1) * e + k] = v[k];
    }
    d_elvtx.Upload(elvtx);
    const size_t nwords = (ne + 31) / 32;
    d_neg.Resize(nwords); d_pos.Resize(nwords); d_if.Resize(nwords); d_ratio.Resize(ne);
  }
  ~GpuCutInformation() override { if (ctx) xgpu_destroy(ctx); }

  void Update(shared_ptr<CoefficientFunction> lset, int subdivlvl, int time_order, LocalHeap& lh) override
  {
    auto [cf, gf] = CF2GFForStraightCutRule(lset, subdivlvl);          // cutint/lsetintdomain.cpp:155-172
    if (!gf || subdivlvl != 0 || time_order >= 0)
      throw Exception("GpuCutInformation: P1 level sets, subdivlvl = 0, no time slab (xgpu_classify_st covers slabs)");
    auto phi = gf->GetVector().FVDouble();
    if (phi.Size() != ma->GetNV()) throw Exception("GpuCutInformation: the level set must have one value per vertex (P1)");
    d_phi.Upload(phi.Data(), phi.Size());                              // the only per-update host-to-device copy
    xgpu_adapter::Check(ctx, xgpu_classify(ctx, dim, (int64_t)ne, d_elvtx, d_phi, d_neg, d_pos, d_if, d_ratio));
    xgpu_adapter::CheckGeometry(ctx);
    // refill the SAME BitArray objects (masks handed out earlier stay valid): NGSolve's BitArray storage (bit i = byte
    // i >> 3, bit i & 7) is byte-identical to the uint32 word layout on a little-endian host
    const size_t nbytes = (ne + 7) / 8;
    d_neg.Download((uint32_t*)elements_of_domain_type[VOL][CDOM_NEG]->Data(), (nbytes + 3) / 4);
    d_pos.Download((uint32_t*)elements_of_domain_type[VOL][CDOM_POS]->Data(), (nbytes + 3) / 4);
    d_if.Download((uint32_t*)elements_of_domain_type[VOL][CDOM_IF]->Data(), (nbytes + 3) / 4);
    d_ratio.Download(cut_ratio_of_element[VOL]->FVDouble().Data(), ne);
    for (auto [comb, a, b] : {std::tuple{CDOM_UNCUT, CDOM_NEG, CDOM_POS}, std::tuple{CDOM_HASNEG, CDOM_NEG, CDOM_IF},
                              std::tuple{CDOM_HASPOS, CDOM_POS, CDOM_IF}}) {
      *elements_of_domain_type[VOL][comb] = *elements_of_domain_type[VOL][a];
      elements_of_domain_type[VOL][comb]->Or(*elements_of_domain_type[VOL][b]);
    }
    // BND elements (cutinfo.cpp:82-125): the same call with dim - 1 on the boundary-element vertex table
  }
};
}  // namespace ngcomp
