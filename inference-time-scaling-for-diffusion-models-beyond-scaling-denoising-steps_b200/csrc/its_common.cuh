// Shared helpers for libits_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/its_b200.h"

namespace its {

// Per-thread error text returned by its_last_error_string().
char* err_buf();
int set_error(int code, const char* fmt, ...);

#define ITS_CHECK_CUDA(expr)                                                   \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess)                                                     \
      return ::its::set_error(ITS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,    \
                              cudaGetErrorString(_e), __FILE__, __LINE__);     \
  } while (0)

#define ITS_REQUIRE(cond, ...)                                                 \
  do {                                                                         \
    if (!(cond)) return ::its::set_error(ITS_ERR_INVALID, __VA_ARGS__);        \
  } while (0)

// Launch-error check that is safe under stream capture (no sync).
#define ITS_CHECK_LAUNCH() ITS_CHECK_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute acts on the current device only: the "already configured" caches of the launchers are
// kept per device (one process normally drives one GPU, but a model on a second device must not inherit the
// first device's flag).
struct PerDeviceBytes {
  size_t v[64] = {};
  // true when the kernel attribute must be (re)set to hold `bytes` of dynamic shared memory on the current device
  bool need(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 63;
    if (bytes <= v[dev]) return false;
    v[dev] = bytes;
    return true;
  }
};

// Programmatic dependent launch: every kernel of the library starts with pdl_prologue()
// (release the next launch, then wait for the previous grid's memory to be visible), so
// consecutive launches on a stream overlap their launch latency / set-up with the tail of
// the predecessor — when the launch carries the attribute (its_set_pdl / ITS_PDL; by default only the
// tap-GEMM launches do, see the measurements in ddpm_step.cu).
bool pdl_enabled(int kind = 0);   // kind: 0 = small kernel, 1 = tap-GEMM, 2 = GroupNorm-apply
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

#define ITS_LAUNCH_KIND(kind_, kernel_, grid_, block_, smem_, stream_, ...)                      \
  do {                                                                                         \
    cudaLaunchConfig_t _cfg = {};                                                              \
    _cfg.gridDim = (grid_);                                                                    \
    _cfg.blockDim = (block_);                                                                  \
    _cfg.dynamicSmemBytes = (smem_);                                                           \
    _cfg.stream = (stream_);                                                                   \
    cudaLaunchAttribute _at[1];                                                                \
    _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
    _at[0].val.programmaticStreamSerializationAllowed = 1;                                     \
    _cfg.attrs = _at;                                                                          \
    _cfg.numAttrs = ::its::pdl_enabled(kind_) ? 1 : 0;                                         \
    ITS_CHECK_CUDA(cudaLaunchKernelEx(&_cfg, kernel_, __VA_ARGS__));                           \
  } while (0)

#define ITS_LAUNCH(kernel_, grid_, block_, smem_, stream_, ...)                                   \
  do {                                                                                         \
    cudaLaunchConfig_t _cfg = {};                                                              \
    _cfg.gridDim = (grid_);                                                                    \
    _cfg.blockDim = (block_);                                                                  \
    _cfg.dynamicSmemBytes = (smem_);                                                           \
    _cfg.stream = (stream_);                                                                   \
    cudaLaunchAttribute _at[1];                                                                \
    _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
    _at[0].val.programmaticStreamSerializationAllowed = 1;                                     \
    _cfg.attrs = _at;                                                                          \
    _cfg.numAttrs = ::its::pdl_enabled() ? 1 : 0;                                              \
    ITS_CHECK_CUDA(cudaLaunchKernelEx(&_cfg, kernel_, __VA_ARGS__));                           \
  } while (0)

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
// Swish with the fast-math reciprocal instead of an IEEE division: x * rcp(1 + exp(-x)), relative
// error a few fp32 ulps (tanh.approx would be cheaper but its 2^-11 error is visible next to bf16's
// 2^-9 rounding in the 2e-2 sample tolerance once guidance amplifies it)
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// Swish through ONE special-function operation: x * sigmoid(x) = h + h * tanh(h) with h = x / 2.  tanh.approx.f32
// has a relative error of 2^-11, so the result carries an absolute error of at most |x| * 2.4e-4 — the size of
// the IEEE fp16 rounding (2^-11 relative) the GroupNorm outputs receive when they are stored anyway.  The
// GroupNorm-apply kernels are bound by the special-function unit (16 results per clock per SM against 128 FMAs):
// exp2 + reciprocal is two operations per element, this is one (ITS_SWISH=exact restores the former).
__device__ __forceinline__ float silu_tanh_half(float h) {     // h = x / 2
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
bool swish_tanh_enabled();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 bf16 <-> 8 floats through one 16-byte access.
struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// 8 floats -> 8 IEEE halves in the same 16-byte container (GroupNorm outputs, see its_src_t.fp16)
__device__ __forceinline__ bf16x8 pack8_half(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    p.v[i] = *reinterpret_cast<const __nv_bfloat162*>(&h);
  }
  return p;
}

// 8 IEEE halves -> 8 floats, and the format-dispatching pair used where a tensor may hold either 16-bit format
__device__ __forceinline__ void unpack8_half(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&p.v[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack8_fmt(const bf16x8& p, float* f, bool half) {
  if (half) unpack8_half(p, f); else unpack8(p, f);
}
__device__ __forceinline__ bf16x8 pack8_fmt(const float* f, bool half) { return half ? pack8_half(f) : pack8(f); }
// two packed 16-bit values of either format -> floats
__device__ __forceinline__ void decode2_fmt(uint32_t w, bool half, float& a, float& b) {
  if (half) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
    a = t.x; b = t.y;
  } else {
    a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
  }
}

// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the DDPM
// step and the candidate generators share.  The host/oracle restatement is
// oracle/philox.py.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c[0];
    uint64_t p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0;
    uint32_t n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += W0; k1 += W1;
  }
}

// Four N(0,1) draws for (seed, candidate, tag, quad index): Box-Muller on two
// pairs of 32-bit uniforms mapped to (0,1].
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t cand, uint32_t tag,
                                               uint32_t quad, float z[4]) {
  uint32_t c[4] = {quad, (uint32_t)cand, tag, (uint32_t)(cand >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float inv32 = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float u1 = ((float)c[2 * i] + 1.0f) * inv32;      // (0,1]
    float u2 = (float)c[2 * i + 1] * inv32;           // [0,1)
    float r = sqrtf(-2.0f * __logf(u1));
    float s, co;
    __sincosf(6.283185307179586f * u2, &s, &co);
    z[2 * i] = r * co;
    z[2 * i + 1] = r * s;
  }
}

}  // namespace its
