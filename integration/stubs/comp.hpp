// integration/stubs/comp.hpp — NOT NGSolve.  A minimal stand-in for the handful of NGSolve declarations the adapter sources
// in integration/ touch, written from NGSolve's public headers as remembered, so that `g++ -fsyntax-only` can check the
// adapters' own code (types, argument orders of the xgpu_* calls) in an image without NGSolve.  Building against the real
// <comp.hpp> is the maintainer's step; nothing here is linked or shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace ngcore {
using std::shared_ptr;
using std::make_shared;
using std::string;
class Exception : public std::exception {
  std::string m;
public:
  Exception(const std::string& s) : m(s) {}
  Exception(const char* s) : m(s) {}
  const char* what() const noexcept override { return m.c_str(); }
};
template <class T> std::string ToString(const T& t) { return std::to_string(t); }
template <class T> class FlatArray {
protected:
  size_t n = 0; T* p = nullptr;
public:
  FlatArray() = default;
  FlatArray(size_t an, T* ap) : n(an), p(ap) {}
  size_t Size() const { return n; }
  T* Data() const { return p; }
  T& operator[](size_t i) const { return p[i]; }
  T* begin() const { return p; }
  T* end() const { return p + n; }
};
template <class T> class Array : public FlatArray<T> {
  std::vector<T> st;
public:
  Array() = default;
  explicit Array(size_t an) : st(an) { this->n = an; this->p = st.data(); }
  void SetSize(size_t an) { st.resize(an); this->n = an; this->p = st.data(); }
};
class BitArray {
  size_t n = 0; std::vector<unsigned char> d;
public:
  BitArray() = default;
  explicit BitArray(size_t an) : n(an), d((an + 7) / 8) {}
  size_t Size() const { return n; }
  unsigned char* Data() { return d.data(); }
  bool Test(size_t i) const { return (d[i >> 3] >> (i & 7)) & 1; }
  BitArray& Or(const BitArray& o) { for (size_t i = 0; i < d.size(); i++) d[i] |= o.d[i]; return *this; }
};
class LocalHeap { public: explicit LocalHeap(size_t, const char* = "lh") {} };
struct IntRange { size_t a, b; struct It { size_t i; size_t operator*() const { return i; } It& operator++() { ++i; return *this; } bool operator!=(It o) const { return i != o.i; } }; It begin() const { return {a}; } It end() const { return {b}; } };
inline IntRange Range(size_t n) { return {0, n}; }
}  // namespace ngcore

namespace ngbla {
template <class T = double> class FlatVector {
  size_t n = 0; T* p = nullptr;
public:
  FlatVector() = default;
  FlatVector(size_t an, T* ap) : n(an), p(ap) {}
  size_t Size() const { return n; }
  T* Data() const { return p; }
  T& operator()(size_t i) const { return p[i]; }
};
template <int N, class T = double> struct Vec { T v[N]; T& operator()(int i) { return v[i]; } const T& operator()(int i) const { return v[i]; } };
}  // namespace ngbla

namespace ngfem {
enum VorB { VOL = 0, BND = 1, BBND = 2 };
class CoefficientFunction { public: virtual ~CoefficientFunction() = default; };
}  // namespace ngfem

namespace ngla {
class BaseVector { public: virtual ~BaseVector() = default; virtual ngbla::FlatVector<double> FVDouble() const = 0; };
template <class TM> class SparseMatrix {
public:
  virtual ~SparseMatrix() = default;
  size_t Height() const;
  size_t NZE() const;
  ngcore::FlatArray<size_t> GetFirstArray() const;       // row pointers ("firsti")
  ngcore::FlatArray<int> GetColIndices() const;          // all column indices ("colnr")
  ngcore::FlatArray<TM> GetValues() const;               // all values, row after row
};
}  // namespace ngla

namespace ngcomp {
using namespace ngcore;
using namespace ngbla;
using namespace ngfem;
using namespace ngla;
struct ElementId { VorB vb; size_t nr; ElementId(VorB avb, size_t anr) : vb(avb), nr(anr) {} size_t Nr() const { return nr; } };
class Ngs_Element : public ElementId {
public:
  using ElementId::ElementId;
  FlatArray<int> Vertices() const;
};
class MeshAccess {
public:
  int GetDimension() const;
  size_t GetNE(VorB vb = VOL) const;
  size_t GetNV() const;
  Ngs_Element GetElement(ElementId ei) const;
  template <int D> void GetPoint(size_t pi, Vec<D>& p) const;
};
class FESpace {
public:
  virtual ~FESpace() = default;
  shared_ptr<MeshAccess> GetMeshAccess() const;
  size_t GetNDof() const;
  int GetOrder() const;
  void GetDofNrs(ElementId ei, Array<int>& dnums) const;
};
class GridFunction : public CoefficientFunction {
public:
  shared_ptr<FESpace> GetFESpace() const;
  BaseVector& GetVector() const;
};
template <class TM, class TV = TM> class T_BilinearForm {
protected:
  shared_ptr<FESpace> fespace;
  shared_ptr<SparseMatrix<TM>> mat;       // allocated by AllocateMatrix() from GetGraph()
public:
  virtual ~T_BilinearForm() = default;
  virtual void DoAssemble(LocalHeap& lh) = 0;
};
}  // namespace ngcomp
