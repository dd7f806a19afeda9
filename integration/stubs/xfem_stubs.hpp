// integration/stubs/xfem_stubs.hpp — NOT ngsxfem.  The few declarations of ngsxfem's own headers (xfem/cutinfo.hpp:17-160,
// utils/restrictedblf.hpp:14-60, utils/ngsxstd.hpp:16-40, cutint/lsetintdomain.hpp) that the adapters use, for the syntax check.
#pragma once
#include <comp.hpp>
#include <tuple>

namespace ngcomp {
enum DOMAIN_TYPE { NEG = 0, POS = 1, IF = 2 };
enum COMBINED_DOMAIN_TYPE { CDOM_NO = 0, CDOM_NEG = 1, CDOM_POS = 2, CDOM_UNCUT = 3, CDOM_IF = 4, CDOM_HASNEG = 5, CDOM_HASPOS = 6, CDOM_ANY = 7 };

std::tuple<shared_ptr<CoefficientFunction>, shared_ptr<GridFunction>> CF2GFForStraightCutRule(shared_ptr<CoefficientFunction> cf, int subdivlvl);

class CutInformation {   // xfem/cutinfo.hpp:17-130
protected:
  shared_ptr<MeshAccess> ma;
  shared_ptr<BaseVector> cut_ratio_of_element[2];
  shared_ptr<BitArray> elements_of_domain_type[2][8];
public:
  CutInformation(shared_ptr<MeshAccess> ama) : ma(ama) {}
  virtual ~CutInformation() = default;
  virtual void Update(shared_ptr<CoefficientFunction> lset, int subdivlvl, int time_order, LocalHeap& lh) = 0;
};

template <class TM, class TV = TM>
class RestrictedBilinearForm : public T_BilinearForm<TM, TV> {   // utils/restrictedblf.hpp:14-60
protected:
  shared_ptr<BitArray> el_restriction, fac_restriction;
};
}  // namespace ngcomp
