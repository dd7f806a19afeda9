// integration/gpu_common.hpp — small helpers of the NGSolve-side adapters: RAII device buffers over the CUDA
// runtime and the error bridge  xgpu status code -> ngcore::Exception  (the reference throws at the same places).
#pragma once
#include <cuda_runtime_api.h>
#include <comp.hpp>
#include "xgpu.h"

namespace xgpu_adapter {
using namespace ngcomp;

inline void Check(xgpu_ctx* ctx, int rc)
{
  if (rc != XGPU_OK) throw Exception(std::string("xgpu: ") + xgpu_last_error(ctx));
}

inline void CheckGeometry(xgpu_ctx* ctx)
{
  int code = 0; int64_t el = -1;
  Check(ctx, xgpu_get_status(ctx, &code, &el));
  if (code) throw Exception("cut geometry error (code " + ToString(code) + ") on element " + ToString((long long)el));   // straightcutrule.cpp:45-49, 144-149, 190-194
}

template <class T>
class DeviceBuffer {
  T* p = nullptr; size_t n = 0;
public:
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { if (p) cudaFree(p); }
  void Resize(size_t an)
  {
    if (an <= n) return;
    if (p) cudaFree(p);
    if (cudaMalloc((void**)&p, an * sizeof(T)) != cudaSuccess) throw Exception("xgpu adapter: out of device memory");
    n = an;
  }
  size_t Size() const { return n; }
  operator T*() const { return p; }
  void Upload(const T* src, size_t cnt) { Resize(cnt); if (cudaMemcpy(p, src, cnt * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) throw Exception("xgpu adapter: H2D copy failed"); }
  void Upload(FlatArray<T> a) { Upload(a.Data(), a.Size()); }
  void Download(T* dst, size_t cnt) const { if (cudaMemcpy(dst, p, cnt * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) throw Exception("xgpu adapter: D2H copy failed"); }
};
}  // namespace xgpu_adapter
