// Shared host/device helpers for libdlpm_b200.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/dlpm_b200.h"

namespace dlpm {

void set_error(const char* fmt, ...);

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DLPM_ERR_CUDA;
}

#define DLPM_CHECK_LAUNCH(what)                                  \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::dlpm::cuda_fail(e__, what); \
  } while (0)

#define DLPM_REQUIRE(cond, msg)          \
  do {                                   \
    if (!(cond)) {                       \
      ::dlpm::set_error("%s", msg);      \
      return DLPM_ERR_ARG;               \
    }                                    \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * max_waves * (2048 / threads);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// Programmatic dependent launch (PDL): every kernel of the per-step chain is launched with the programmatic-stream-
// serialization attribute; it lets its successor start launching right away (pdl_launch_dependents) and blocks
// (pdl_wait) until its predecessor has fully completed before it touches global memory.  The successor's launch
// latency and prologue (barrier init, TMEM allocation, descriptor prefetch, weight staging) overlap the predecessor's
// tail.  Both instructions are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();
void pdl_set_enabled(bool on);

// cudaLaunchKernelEx wrapper: optional thread-block cluster (x dimension) and the PDL attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Division by a launch-constant via multiply-high (valid for 0 <= n < 2^31): the streaming kernels index
// (sample, position) from a flat quad index; a 64-bit hardware-emulated divide there costs more than the memory traffic.
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t div) : d(div), mul(0), shr(0) {
    if (div > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < div) ++lg;
      const unsigned p = 31 + lg;
      mul = (uint32_t)(((1ull << p) + div - 1) / div);
      shr = p - 32;
    }
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

// streaming 128-bit accesses (data touched once: bypass L1 allocation)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_rw(const float4* p) {  // plain (the location is rewritten by the same thread)
  return *p;
}

__device__ __forceinline__ float4 bf16x4_to_float4(uint2 raw) {
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

}  // namespace dlpm
