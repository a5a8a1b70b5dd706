// GroupNorm(32) [+ Swish] over NHWC bf16, with the skip-connection concat folded
// in (two sources, one output).  HBM/L2-bound: one 16-byte load per 8 channels in
// each of the two passes, one 16-byte store.  Deterministic (no float atomics):
// pass 1 writes per-(image, chunk, group) partial sums, pass 2 reduces them in a
// fixed order in double and applies scale/shift(+Swish).
// Reference: nn.GroupNorm(32, C) + Swish at Model.py:170-173,186-190,132,257-259;
// concat at Model.py:279-280.
#include "its_common.cuh"
#include <cooperative_groups.h>

namespace its {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_VEC = 128;  // C <= 1024


struct GnArgs {
  const __nv_bfloat16* src0;
  const __nv_bfloat16* src1;
  __nv_bfloat16* out;
  const float* gamma;
  const float* beta;
  float* partials;
  int C0, C1, C, HW, groups, chunks, silu, out_fp16;
  int src0_fp16, src1_fp16;     // 16-bit format of the sources
  float eps;
};

__device__ __forceinline__ bool gn_half(const GnArgs& a, int c0) { return (c0 < a.C0 ? a.src0_fp16 : a.src1_fp16) != 0; }

__device__ __forceinline__ bf16x8 gn_load(const GnArgs& a, long long pix, int c0) {
  // pix = global pixel index (image*HW + p); c0 = first channel of the vector
  if (c0 < a.C0)
    return *reinterpret_cast<const bf16x8*>(a.src0 + pix * a.C0 + c0);
  return *reinterpret_cast<const bf16x8*>(a.src1 + pix * a.C1 + (c0 - a.C0));
}

__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const GnArgs a) {
  pdl_prologue();
  __shared__ float s_sum[GN_THREADS * 8];
  __shared__ float s_sq[GN_THREADS * 8];
  const int nvec = a.C / 8;
  const int prow = GN_THREADS / nvec;  // pixel lanes per CTA (>= 2 for C <= 1024)
  const int tid = threadIdx.x;
  const int cv = tid % nvec, pl = tid / nvec;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  if (pl < prow) {
    for (int p = p0 + pl; p < p1; p += prow) {
      float f[8];
      unpack8_fmt(gn_load(a, (long long)img * a.HW + p, cv * 8), f, gn_half(a, cv * 8));
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum[i] += f[i]; sq[i] = fmaf(f[i], f[i], sq[i]); }
    }
    // layout [pl][channel]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s_sum[pl * a.C + cv * 8 + i] = sum[i];
      s_sq[pl * a.C + cv * 8 + i] = sq[i];
    }
  }
  __syncthreads();
  // one warp per group (strided), fixed summation order -> deterministic
  const int warp = tid >> 5, lane = tid & 31;
  const int cg = a.C / a.groups;
  for (int g = warp; g < a.groups; g += GN_THREADS / 32) {
    float ps = 0.f, pq = 0.f;
    const int n = cg * prow;
    for (int e = lane; e < n; e += 32) {
      const int r = e / cg, c = g * cg + (e - r * cg);
      ps += s_sum[r * a.C + c];
      pq += s_sq[r * a.C + c];
    }
    ps = warp_sum(ps);
    pq = warp_sum(pq);
    if (lane == 0) {
      float* dst = a.partials + (((long long)img * a.chunks + chunk) * a.groups + g) * 2;
      dst[0] = ps;
      dst[1] = pq;
    }
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const GnArgs a) {
  pdl_prologue();
  __shared__ float s_mean[64], s_rstd[64];
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int cg = a.C / a.groups;
  if (tid < a.groups) {
    double s = 0.0, q = 0.0;
    const float* src = a.partials + ((long long)img * a.chunks * a.groups + tid) * 2;
    for (int c = 0; c < a.chunks; ++c) {
      s += (double)src[(long long)c * a.groups * 2];
      q += (double)src[(long long)c * a.groups * 2 + 1];
    }
    const double n = (double)cg * (double)a.HW;
    const double mean = s / n;
    double var = q / n - mean * mean;  // biased, like nn.GroupNorm
    if (var < 0.0) var = 0.0;
    s_mean[tid] = (float)mean;
    s_rstd[tid] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  __syncthreads();
  const int nvec = a.C / 8;
  const int prow = GN_THREADS / nvec;
  const int cv = tid % nvec, pl = tid / nvec;
  if (pl >= prow) return;
  float scale[8], shift[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    const int g = c / cg;
    const float sc = s_rstd[g] * a.gamma[c];
    scale[i] = sc;
    shift[i] = a.beta[c] - s_mean[g] * sc;
  }
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  for (int p = p0 + pl; p < p1; p += prow) {
    const long long pix = (long long)img * a.HW + p;
    float f[8];
    unpack8_fmt(gn_load(a, pix, cv * 8), f, gn_half(a, cv * 8));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(f[i], scale[i], shift[i]);
      f[i] = a.silu ? silu_f(v) : v;
    }
    *reinterpret_cast<bf16x8*>(a.out + pix * a.C + cv * 8) = a.out_fp16 ? pack8_half(f) : pack8(f);
  }
}

// Single-launch variant: the `chunks` CTAs of one image form a thread-block
// cluster, exchange their per-group partial sums through distributed shared memory
// (fixed rank order -> deterministic) and apply the normalisation to their chunk,
// which they kept in shared memory (CACHE) — one global read + one write per element.
template <bool CACHE>
__global__ void __launch_bounds__(GN_THREADS) gn_cluster_kernel(const GnArgs a) {
  pdl_prologue();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char gn_dyn[];
  __shared__ float s_sum[GN_THREADS * 8];
  __shared__ float s_sq[GN_THREADS * 8];
  __shared__ float s_part[64 * 2];
  __shared__ float s_mean[64], s_rstd[64];
  __nv_bfloat16* cache = reinterpret_cast<__nv_bfloat16*>(gn_dyn);
  const int nvec = a.C / 8;
  const int prow = GN_THREADS / nvec;
  const int tid = threadIdx.x;
  const int cv = tid % nvec, pl = tid / nvec;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  const int cg_ch = a.C / a.groups;
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sum[i] = sq[i] = 0.f;
  if (pl < prow) {
    constexpr int U = 8;   // independent 16-byte loads in flight per thread
    for (int pb = p0 + pl; pb < p1; pb += prow * U) {
      bf16x8 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * prow;
        if (p < p1) raw[u] = gn_load(a, (long long)img * a.HW + p, cv * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * prow;
        if (p < p1) {
          if (CACHE) *reinterpret_cast<bf16x8*>(cache + (long long)(p - p0) * a.C + cv * 8) = raw[u];
          float f[8];
          unpack8_fmt(raw[u], f, gn_half(a, cv * 8));
#pragma unroll
          for (int i = 0; i < 8; ++i) { sum[i] += f[i]; sq[i] = fmaf(f[i], f[i], sq[i]); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s_sum[pl * a.C + cv * 8 + i] = sum[i];
      s_sq[pl * a.C + cv * 8 + i] = sq[i];
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int g = warp; g < a.groups; g += GN_THREADS / 32) {
    float ps = 0.f, pq = 0.f;
    const int n = cg_ch * prow;
    for (int e = lane; e < n; e += 32) {
      const int r = e / cg_ch, c = g * cg_ch + (e - r * cg_ch);
      ps += s_sum[r * a.C + c];
      pq += s_sq[r * a.C + c];
    }
    ps = warp_sum(ps);
    pq = warp_sum(pq);
    if (lane == 0) { s_part[2 * g] = ps; s_part[2 * g + 1] = pq; }
  }
  cluster.sync();                                  // every CTA's s_part is published
  if (tid < a.groups) {
    double s = 0.0, q = 0.0;
    for (int r = 0; r < a.chunks; ++r) {           // fixed order: identical on all CTAs
      const float* remote = cluster.map_shared_rank(s_part, r);
      s += (double)remote[2 * tid];
      q += (double)remote[2 * tid + 1];
    }
    const double n = (double)cg_ch * (double)a.HW;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = (float)mean;
    s_rstd[tid] = (float)(1.0 / sqrt(var + (double)a.eps));
  }
  cluster.sync();                                  // remote reads done before any CTA may exit
  if (pl >= prow) return;
  float scale[8], shift[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    const int g = c / cg_ch;
    const float sc = s_rstd[g] * a.gamma[c];
    scale[i] = sc;
    shift[i] = a.beta[c] - s_mean[g] * sc;
  }
  for (int p = p0 + pl; p < p1; p += prow) {
    const long long pix = (long long)img * a.HW + p;
    float f[8];
    if (CACHE)
      unpack8_fmt(*reinterpret_cast<const bf16x8*>(cache + (long long)(p - p0) * a.C + cv * 8), f, gn_half(a, cv * 8));
    else
      unpack8_fmt(gn_load(a, pix, cv * 8), f, gn_half(a, cv * 8));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = fmaf(f[i], scale[i], shift[i]);
      f[i] = a.silu ? silu_f(v) : v;
    }
    *reinterpret_cast<bf16x8*>(a.out + pix * a.C + cv * 8) = a.out_fp16 ? pack8_half(f) : pack8(f);
  }
}


// ---- apply-only variant: statistics come from the producing tap-GEMM ----------------
// The persistent tap-GEMM (conv_persist_sm100.cu) leaves, for the tensor it stores, the sum
// and the sum of squares of every (image, 4-channel chunk, tile) in a partial-sum array
// [n_img][parts][C/4] of float2.  This kernel reduces them per group in a fixed order in
// double (tpg lanes per group, xor-shuffle tree), then streams the tensor once:
// 16-byte load -> scale/shift(+Swish) -> 16-byte store.
struct GnStatArgs {
  const __nv_bfloat16* src0;
  const __nv_bfloat16* src1;
  __nv_bfloat16* out;
  const float* gamma;
  const float* beta;
  const float2* st0;
  const float2* st1;
  int C0, C1, C, HW, groups, parts0, parts1, silu, chunks, out_fp16;
  int src0_fp16, src1_fp16;     // 16-bit format of the sources (raw feature maps: bf16 unless the fp16 residual stream is on)
  float eps;
  int late_trigger;             // 1: do not release the dependent launch at kernel start (ITS_PDL=4 experiment)
};

// FMT: 16-bit format of the sources at compile time — 0 = bf16, 1 = IEEE fp16, 2 = per source (run-time flags)
template <int FMT, bool TANH>
__global__ void __launch_bounds__(GN_THREADS, 4) gn_apply_stats_kernel(const GnStatArgs a) {
  __shared__ float s_mean[64], s_rstd[64];
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int cg = a.C / a.groups;         // channels per group (multiple of 4)
  const int cg4 = cg >> 2;
  const int tpg = GN_THREADS / a.groups; // lanes per group: power of two, <= 32
  const int nvec = a.C / 8;
  const int prow = GN_THREADS / nvec;
  const int cv = tid % nvec, pl = tid / nvec;
  const bool active = pl < prow;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  const bool from0 = cv * 8 < a.C0;
  const __nv_bfloat16* sp = from0 ? a.src0 + cv * 8 : a.src1 + (cv * 8 - a.C0);
  const int spitch = from0 ? a.C0 : a.C1;
  const bool src_half = FMT == 2 ? ((from0 ? a.src0_fp16 : a.src1_fp16) != 0) : (FMT == 1);
  constexpr int U = 4;                   // independent 16-byte loads in flight per thread
  if (a.late_trigger) asm volatile("griddepcontrol.wait;" ::: "memory"); else pdl_prologue();
  // first batch of activations and the affine parameters are requested before the
  // statistics chain below, so the two global round trips overlap
  bf16x8 raw[U];
  float gam[8], bet[8];
  if (active) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + pl + u * prow;
      if (p < p1) raw[u] = *reinterpret_cast<const bf16x8*>(sp + ((long long)img * a.HW + p) * spitch);
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + cv * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + cv * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + cv * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + cv * 8) + 1);
    gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
    bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
  }
  {
    const int g = tid / tpg, l = tid - g * tpg;
    double s = 0.0, q = 0.0;
    if (g < a.groups) {
      for (int kk = 0; kk < cg4; ++kk) {
        int k = g * cg4 + kk;            // 4-channel chunk in the concatenated channel space
        const float2* st;
        int parts, nchunk;
        if (k < (a.C0 >> 2)) { st = a.st0; parts = a.parts0; nchunk = a.C0 >> 2; }
        else { st = a.st1; parts = a.parts1; nchunk = a.C1 >> 2; k -= (a.C0 >> 2); }
        for (int part = l; part < parts; part += tpg) {
          const float2 v = __ldg(st + ((long long)img * parts + part) * nchunk + k);
          s += (double)v.x;
          q += (double)v.y;
        }
      }
    }
    for (int o = tpg >> 1; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (g < a.groups && l == 0) {
      const double n = (double)cg * (double)a.HW;
      const double mean = s / n;
      double var = q / n - mean * mean;  // biased, like nn.GroupNorm
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
  }
  __syncthreads();
  if (!active) return;
  float scale[8], shift[8], nscale[8], nshift[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (cv * 8 + i) / cg;
    scale[i] = s_rstd[g] * gam[i];
    shift[i] = bet[i] - s_mean[g] * scale[i];
    // second affine map of the input: the exponent of the exact form, or half the normalised value for tanh
    nscale[i] = (TANH ? 0.5f : -1.4426950408889634f) * scale[i];
    nshift[i] = (TANH ? 0.5f : -1.4426950408889634f) * shift[i];
  }
  for (int pb = p0 + pl; pb < p1; pb += prow * U) {
    if (pb != p0 + pl) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * prow;
        if (p < p1) raw[u] = *reinterpret_cast<const bf16x8*>(sp + ((long long)img * a.HW + p) * spitch);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * prow;
      if (p < p1) {
        float f[8];
        unpack8_fmt(raw[u], f, src_half);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (a.silu && TANH) {
            f[i] = silu_tanh_half(fmaf(f[i], nscale[i], nshift[i]));   // h + h tanh(h), h = v / 2
          } else if (a.silu) {
            // v * rcp(1 + 2^(-v*log2e)); the exponent comes straight from the input (one FMA)
            const float v = fmaf(f[i], scale[i], shift[i]);
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(f[i], nscale[i], nshift[i])));
            f[i] = __fdividef(v, 1.0f + e);
          } else {
            f[i] = fmaf(f[i], scale[i], shift[i]);
          }
        }
        *reinterpret_cast<bf16x8*>(a.out + ((long long)img * a.HW + p) * a.C + cv * 8) = a.out_fp16 ? pack8_half(f) : pack8(f);
      }
    }
  }
}

template <bool CACHE>
static int launch_gn_cluster(const GnArgs& a, int n_img, size_t dyn_bytes, cudaStream_t stream) {
  auto kern = gn_cluster_kernel<CACHE>;
  static PerDeviceBytes configured;
  if (dyn_bytes > 0 && configured.need(dyn_bytes))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.chunks, n_img, 1);
  cfg.blockDim = dim3(GN_THREADS, 1, 1);
  cfg.dynamicSmemBytes = dyn_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.chunks;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  ITS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return ITS_OK;
}

}  // namespace its

extern "C" int its_group_norm(void* out, const void* src0, int32_t C0, const void* src1, int32_t C1,
                              const float* gamma, const float* beta, int32_t n_img, int32_t HW,
                              int32_t groups, float eps, int32_t silu, float* partials,
                              int32_t chunks, int32_t out_fp16, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && src0 && gamma && beta && partials, "its_group_norm: null pointer");
  ITS_REQUIRE(C1 == 0 || src1 != nullptr, "its_group_norm: C1 > 0 needs src1");
  const int C = C0 + C1;
  ITS_REQUIRE(C0 > 0 && C0 % 8 == 0 && C1 % 8 == 0 && C <= 8 * GN_MAX_VEC,
              "its_group_norm: channels (%d+%d) must be multiples of 8 and <= %d", C0, C1, 8 * GN_MAX_VEC);
  ITS_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "its_group_norm: C=%d not divisible by groups=%d", C, groups);
  ITS_REQUIRE(n_img > 0 && HW > 0 && chunks > 0 && chunks <= HW, "its_group_norm: bad n_img/HW/chunks");
  GnArgs a;
  a.src0 = static_cast<const __nv_bfloat16*>(src0);
  a.src1 = static_cast<const __nv_bfloat16*>(src1);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.gamma = gamma; a.beta = beta; a.partials = partials;
  a.C0 = C0; a.C1 = C1; a.C = C; a.HW = HW; a.groups = groups; a.chunks = chunks; a.silu = silu;
  a.out_fp16 = out_fp16 & 1; a.src0_fp16 = (out_fp16 >> 1) & 1; a.src1_fp16 = (out_fp16 >> 2) & 1;
  a.eps = eps;
  if (chunks <= 8) {
    // one launch: cluster of `chunks` CTAs per image
    const int per = (HW + chunks - 1) / chunks;
    const size_t cache_bytes = (size_t)per * C * 2;
    if (cache_bytes <= 96 * 1024) return launch_gn_cluster<true>(a, n_img, cache_bytes, as_stream(stream));
    return launch_gn_cluster<false>(a, n_img, 0, as_stream(stream));
  }
  dim3 grid(chunks, n_img);
  ITS_LAUNCH(gn_stats_kernel, dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), a);
  ITS_CHECK_LAUNCH();
  ITS_LAUNCH(gn_apply_kernel, dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), a);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_group_norm_apply(void* out, const void* src0, int32_t C0, const float* stats0,
                                    int32_t parts0, const void* src1, int32_t C1, const float* stats1,
                                    int32_t parts1, const float* gamma, const float* beta, int32_t n_img,
                                    int32_t HW, int32_t groups, float eps, int32_t silu, int32_t out_fp16,
                                    void* stream) {
  using namespace its;
  ITS_REQUIRE(out && src0 && stats0 && gamma && beta, "its_group_norm_apply: null pointer");
  ITS_REQUIRE(C1 == 0 || (src1 != nullptr && stats1 != nullptr), "its_group_norm_apply: C1 > 0 needs src1 and stats1");
  const int C = C0 + C1;
  // one thread per 8-channel vector of a pixel row: up to GN_THREADS vectors (the 1024+1024 concatenations of
  // MainCondition.py's default widths)
  ITS_REQUIRE(C0 > 0 && C0 % 8 == 0 && C1 % 8 == 0 && C <= 8 * GN_THREADS,
              "its_group_norm_apply: channels (%d+%d) must be multiples of 8 and <= %d", C0, C1, 8 * GN_THREADS);
  ITS_REQUIRE(groups > 0 && groups <= 64 && (groups & (groups - 1)) == 0 && C % groups == 0 && (C / groups) % 4 == 0,
              "its_group_norm_apply: groups=%d must be a power of two dividing C=%d into multiples of 4 channels", groups, C);
  ITS_REQUIRE(n_img > 0 && HW > 0 && parts0 > 0 && (C1 == 0 || parts1 > 0), "its_group_norm_apply: bad n_img/HW/parts");
  GnStatArgs a;
  a.src0 = static_cast<const __nv_bfloat16*>(src0);
  a.src1 = static_cast<const __nv_bfloat16*>(src1);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.gamma = gamma; a.beta = beta;
  a.st0 = reinterpret_cast<const float2*>(stats0);
  a.st1 = reinterpret_cast<const float2*>(stats1);
  a.C0 = C0; a.C1 = C1; a.C = C; a.HW = HW; a.groups = groups; a.parts0 = parts0; a.parts1 = parts1;
  // out_fp16 carries three flags: bit 0 = output format, bit 1 / bit 2 = source 0 / source 1 hold IEEE fp16
  a.silu = silu; a.eps = eps; a.out_fp16 = out_fp16 & 1; a.src0_fp16 = (out_fp16 >> 1) & 1; a.src1_fp16 = (out_fp16 >> 2) & 1;
  // one wave: at most 4 CTAs per SM are resident (launch bounds), and every CTA repeats the
  // statistics prologue of its image, so the grid is sized to fit the machine exactly once
  long long chunks = ((long long)HW * C * 2 + 32767) / 32768;
  const long long fit = (4LL * 148) / n_img;
  if (chunks > fit) chunks = fit;
  if (chunks < 1) chunks = 1;
  if (chunks > HW) chunks = HW;
  a.chunks = (int)chunks;
  dim3 grid((unsigned)chunks, (unsigned)n_img);
  const int s1 = (C1 > 0) ? a.src1_fp16 : a.src0_fp16;
  const bool th = swish_tanh_enabled();
  a.late_trigger = pdl_enabled(2) && !pdl_enabled(0);
#define ITS_GN_APPLY(FMT_)                                                                                       \
  do {                                                                                                          \
    if (th) { ITS_LAUNCH_KIND(2, (gn_apply_stats_kernel<FMT_, true>), dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), a); } \
    else { ITS_LAUNCH_KIND(2, (gn_apply_stats_kernel<FMT_, false>), dim3(grid), dim3(GN_THREADS), 0, as_stream(stream), a); }   \
  } while (0)
  if (a.src0_fp16 == 0 && s1 == 0) {
    ITS_GN_APPLY(0);
  } else if (a.src0_fp16 == 1 && s1 == 1) {
    ITS_GN_APPLY(1);
  } else {
    ITS_GN_APPLY(2);
  }
#undef ITS_GN_APPLY
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}
