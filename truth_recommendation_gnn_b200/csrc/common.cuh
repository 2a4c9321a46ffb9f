// Shared helpers for the sm_100a kernels behind include/trg_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "../../include/trg_b200.h"

namespace trg {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (persistent grids / workspaces are sized for at most this many CTAs)
constexpr int kMaxDevices = 64;

// SM count of the CURRENT device (cached per device ordinal), capped at kNumSMs: persistent kernels
// launch one CTA per SM and their per-CTA workspaces are sized for kNumSMs.
int grid_sms();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: the "already raised" state is kept
// per device ordinal (a process may drive several GPUs through this ABI) in atomics (host threads may race;
// setting the same value twice is harmless).
struct SmemAttrState {
  std::atomic<int> bytes[kMaxDevices];
};
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, SmemAttrState& st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && st.bytes[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && tracked) st.bytes[dev].store(bytes, std::memory_order_release);
  return e;
}

// A/B and ablation switches exist only in -DTRG_DEBUG builds; the product library never reads the
// environment on a call path.
#ifdef TRG_DEBUG
int debug_env_int(const char* name, int dflt);
#else
inline int debug_env_int(const char*, int dflt) { return dflt; }
#endif

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define TRG_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::trg::set_error(__VA_ARGS__);      \
      return TRG_E_ARG;                   \
    }                                     \
  } while (0)

#define TRG_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::trg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                        \
      return TRG_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

// Launch-error check that does not synchronise (catches bad configs / missing images).
#define TRG_LAUNCH_OK()                                                                    \
  do {                                                                                     \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess) {                                                               \
      ::trg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                          \
      return TRG_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

// ---- 128-bit global access ---------------------------------------------------------------
// Gathered table rows: read-only path, keep in L1/L2 (rows are re-read by other destinations).
__device__ __forceinline__ uint4 ldg_row(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Streaming index data read exactly once: do not allocate in L1.
__device__ __forceinline__ int ldg_stream(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ long long ldg_stream(const long long* p) {
  long long r;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// Bulk L2 prefetch of a whole table row (bytes: multiple of 16, address 16-byte aligned).
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Output rows written once: streaming store.
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- element traits: a 16-byte vector holds VEC elements, accumulated as fp32 -----------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kVec = 4;
  __device__ static __forceinline__ void unpack(uint4 v, float* f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(uint4 v, float* f) {
    // bf16 -> fp32 is a 16-bit shift
    f[0] = __uint_as_float(v.x << 16);
    f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16);
    f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16);
    f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16);
    f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

// ---- vector access of VB = 16 or 8 bytes per lane ------------------------------------------------
// 8-byte lanes let a FULL warp own one 256-byte row (fp32 H=64, bf16 H=128) instead of two half-warps
// walking two rows with divergent control flow.
template <typename T, int VB>
struct Vec;
template <typename T>
struct Vec<T, 16> {
  using Raw = uint4;
  static constexpr int kVec = Elem<T>::kVec;
  __device__ static __forceinline__ Raw load(const void* p) { return ldg_row(p); }
  __device__ static __forceinline__ Raw load_plain(const void* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ static __forceinline__ void unpack(Raw v, float* f) { Elem<T>::unpack(v, f); }
  __device__ static __forceinline__ void store(void* p, const float* f) { stg_stream(p, Elem<T>::pack(f)); }
};
__device__ __forceinline__ uint2 ldg_row8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
template <>
struct Vec<float, 8> {
  using Raw = uint2;
  static constexpr int kVec = 2;
  __device__ static __forceinline__ Raw load(const void* p) { return ldg_row8(p); }
  __device__ static __forceinline__ Raw load_plain(const void* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static __forceinline__ void unpack(Raw v, float* f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
  }
  __device__ static __forceinline__ void store(void* p, const float* f) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])) : "memory");
  }
};
template <>
struct Vec<__nv_bfloat16, 8> {
  using Raw = uint2;
  static constexpr int kVec = 4;
  __device__ static __forceinline__ Raw load(const void* p) { return ldg_row8(p); }
  __device__ static __forceinline__ Raw load_plain(const void* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ static __forceinline__ void unpack(Raw v, float* f) {
    f[0] = __uint_as_float(v.x << 16);
    f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16);
    f[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  __device__ static __forceinline__ void store(void* p, const float* f) {
    const uint32_t a = Elem<__nv_bfloat16>::pack2(f[0], f[1]), b = Elem<__nv_bfloat16>::pack2(f[2], f[3]);
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  }
};

}  // namespace trg
