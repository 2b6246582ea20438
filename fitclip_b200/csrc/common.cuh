// Shared host/device helpers for libfitclip_b200: status codes, error string, bf16 packing, warp reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace fc {

typedef __nv_bfloat16 bf16;

// ---- status codes returned across the C ABI (0 = ok, negative = error; text via fc_last_error) ----
enum Status : int {
  FC_OK = 0,
  FC_ERR_INVALID = -1,   // bad argument (null pointer, shape, alignment)
  FC_ERR_CUDA = -2,      // a CUDA runtime/driver call failed
  FC_ERR_ARCH = -3,      // device is not compute capability 10.x
  FC_ERR_STATE = -4,     // handle used before weights were loaded, etc.
  FC_ERR_NOMEM = -5,
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FC_CUDA(call)                                                     \
  do {                                                                    \
    cudaError_t _e = (call);                                              \
    if (_e != cudaSuccess) return ::fc::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

void note_launch();
#define FC_CHECK_LAUNCH()       \
  do {                          \
    ::fc::note_launch();        \
    FC_CUDA(cudaGetLastError()); \
  } while (0)

#define FC_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::fc::set_error(__VA_ARGS__);      \
      return ::fc::FC_ERR_INVALID;       \
    }                                    \
  } while (0)

int num_sms();

// ---- optional per-launch CUDA-event profiler (fc_profile_start / fc_profile_stop); off by default ----
enum ProfKind : int { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_LAYERNORM = 2, PROF_OTHER = 3 };
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(cudaStream_t s, int kind, int tag, int64_t m, int64_t n, int64_t k, double flops, double bytes);
  ~ProfScope();
};

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 128-bit streaming global accesses (data touched once: keep it out of L1)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// same without .nc: for buffers the kernel may also WRITE (an output that aliases the input row by row)
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

}  // namespace fc
