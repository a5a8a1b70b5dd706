// AttnBlock pieces that are not GEMMs (Model.py:153-158): the row softmax
// between the two tensor-core batched GEMMs, and a one-kernel CUDA-core attention
// for token counts below one 128-row MMA tile (4x4 and 8x8 feature maps).
#include "its_common.cuh"

namespace its {

// One warp per row: fp32 scores -> bf16 probabilities.
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* __restrict__ probs,
                                                           const float* __restrict__ scores,
                                                           long long n_rows, int n_cols) {
  pdl_prologue();
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* s = scores + row * n_cols;
  float m = -INFINITY;
  for (int c = lane; c < n_cols; c += 32) m = fmaxf(m, s[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int c = lane; c < n_cols; c += 32) sum += __expf(s[c] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* p = probs + row * n_cols;
  for (int c = lane; c < n_cols; c += 32) p[c] = __float2bfloat16_rn(__expf(s[c] - m) * inv);
}

// One CTA per (image, query token): for very small maps (N <= 32 tokens) one CTA per image cannot
// keep its warps busy (4 queries per warp pass), so the queries are spread over the grid instead.  qkv [n_img][N][3C] bf16, out [n_img][N][C].
__global__ void __launch_bounds__(256) attention_tiny_kernel(__nv_bfloat16* __restrict__ out,
                                                              const __nv_bfloat16* __restrict__ qkv,
                                                              int N, int C, float scale) {
  pdl_prologue();
  __shared__ float s_p[64];
  const int img = blockIdx.y, qi = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* base = qkv + (long long)img * N * 3 * C;
  const __nv_bfloat16* q = base + (long long)qi * 3 * C;
  const int nv = C / 8;
  for (int j = warp; j < N; j += 8) {
    const __nv_bfloat16* k = base + (long long)j * 3 * C + C;
    float acc = 0.f;
    for (int v = lane; v < nv; v += 32) {
      float fq[8], fk[8];
      unpack8(*reinterpret_cast<const bf16x8*>(q + v * 8), fq);
      unpack8(*reinterpret_cast<const bf16x8*>(k + v * 8), fk);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(fq[i], fk[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_p[j] = acc * scale;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, s_p[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += __expf(s_p[j] - m);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) s_p[j] = __expf(s_p[j] - m) * inv;
  }
  __syncthreads();
  for (int c2 = tid; c2 < C / 2; c2 += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < N; ++j) {
      const float2 vv = __bfloat1622float2(
          *reinterpret_cast<const __nv_bfloat162*>(base + (long long)j * 3 * C + 2 * C + 2 * c2));
      a0 = fmaf(s_p[j], vv.x, a0);
      a1 = fmaf(s_p[j], vv.y, a1);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + ((long long)img * N + qi) * C + 2 * c2) =
        __floats2bfloat162_rn(a0, a1);
  }
}

// One CTA per image: q, k, v rows staged once in shared memory (row pitch C + 8 halfwords, so the
// 16-byte reads of 32 different key rows hit 32 different bank groups), then each warp takes QB
// queries at a time: scores with lane = key (N <= 64: two keys per lane), softmax across the warp,
// and P V with lane = channel pair.  qkv [n_img][N][3C] bf16, out [n_img][N][C].
constexpr int AS_QB = 4;   // (the probability exchange below is written for exactly 4)
__global__ void __launch_bounds__(256) attention_small_kernel(__nv_bfloat16* __restrict__ out,
                                                              const __nv_bfloat16* __restrict__ qkv,
                                                              int N, int C, float scale) {
  extern __shared__ __align__(16) uint8_t as_smem[];
  pdl_prologue();
  const int pitch = C + 8;                                   // halfwords
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(as_smem);
  __nv_bfloat16* sk = sq + (size_t)N * pitch;
  __nv_bfloat16* sv = sk + (size_t)N * pitch;
  float* sp = reinterpret_cast<float*>(sv + (size_t)N * pitch);   // [8 warps][QB][64]
  const int img = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* base = qkv + (long long)img * N * 3 * C;
  const int nv = C / 8, row_v = 3 * nv;
  for (int i = tid; i < N * row_v; i += blockDim.x) {
    const int r = i / row_v, c = i - r * row_v;
    const int which = c / nv, cv = c - which * nv;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * 3 * C) + c);
    *reinterpret_cast<uint4*>((which == 0 ? sq : which == 1 ? sk : sv) + (size_t)r * pitch + cv * 8) = u;
  }
  __syncthreads();
  float* wp = sp + warp * AS_QB * 64;
  for (int q0 = warp * AS_QB; q0 < N; q0 += 8 * AS_QB) {
    float acc[AS_QB][2];
#pragma unroll
    for (int a = 0; a < AS_QB; ++a) acc[a][0] = acc[a][1] = 0.f;
    const int j0 = lane < N ? lane : N - 1, j1 = lane + 32 < N ? lane + 32 : N - 1;
    for (int v = 0; v < nv; ++v) {
      float k0[8], k1[8];
      unpack8(*reinterpret_cast<const bf16x8*>(sk + (size_t)j0 * pitch + v * 8), k0);
      unpack8(*reinterpret_cast<const bf16x8*>(sk + (size_t)j1 * pitch + v * 8), k1);
#pragma unroll
      for (int a = 0; a < AS_QB; ++a) {
        const int qi = q0 + a < N ? q0 + a : N - 1;
        float fq[8];
        unpack8(*reinterpret_cast<const bf16x8*>(sq + (size_t)qi * pitch + v * 8), fq);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[a][0] = fmaf(fq[i], k0[i], acc[a][0]);
          acc[a][1] = fmaf(fq[i], k1[i], acc[a][1]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AS_QB; ++a) {
      const float s0 = lane < N ? acc[a][0] * scale : -INFINITY;
      const float s1 = lane + 32 < N ? acc[a][1] * scale : -INFINITY;
      float m = fmaxf(s0, s1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float e0 = lane < N ? __expf(s0 - m) : 0.f, e1 = lane + 32 < N ? __expf(s1 - m) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      wp[lane * AS_QB + a] = e0 * inv;            // [key][query]: one 16-byte read per key below
      wp[(lane + 32) * AS_QB + a] = e1 * inv;
    }
    __syncwarp();
    for (int c2 = lane; c2 < C / 2; c2 += 32) {
      float o[AS_QB][2];
#pragma unroll
      for (int a = 0; a < AS_QB; ++a) o[a][0] = o[a][1] = 0.f;
      for (int j = 0; j < N; ++j) {
        const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + (size_t)j * pitch + 2 * c2));
        const float4 p4 = *reinterpret_cast<const float4*>(wp + j * AS_QB);
        const float pj[AS_QB] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
        for (int a = 0; a < AS_QB; ++a) {
          o[a][0] = fmaf(pj[a], vv.x, o[a][0]);
          o[a][1] = fmaf(pj[a], vv.y, o[a][1]);
        }
      }
#pragma unroll
      for (int a = 0; a < AS_QB; ++a)
        if (q0 + a < N)
          *reinterpret_cast<__nv_bfloat162*>(out + ((long long)img * N + q0 + a) * C + 2 * c2) =
              __floats2bfloat162_rn(o[a][0], o[a][1]);
    }
    __syncwarp();
  }
}

}  // namespace its

extern "C" int its_softmax_rows(void* probs_bf16, const float* scores, int64_t n_rows,
                                int32_t n_cols, void* stream) {
  ITS_REQUIRE(probs_bf16 && scores && n_rows > 0 && n_cols > 0, "its_softmax_rows: bad arguments");
  const long long blocks = (n_rows * 32 + 255) / 256;
  ITS_REQUIRE(blocks < (1LL << 31), "its_softmax_rows: too many rows");
  ITS_LAUNCH(its::softmax_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, its::as_stream(stream), 
      static_cast<__nv_bfloat16*>(probs_bf16), scores, n_rows, n_cols);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_attention_small(void* out, const void* qkv, int32_t n_img, int32_t N, int32_t C,
                                   float scale, void* stream) {
  ITS_REQUIRE(out && qkv, "its_attention_small: null pointer");
  ITS_REQUIRE(n_img > 0 && n_img <= 65535 && N > 0 && N <= 64 && C > 0 && C % 8 == 0,
              "its_attention_small: unsupported N=%d C=%d n_img=%d", N, C, n_img);
  const size_t smem_need = (size_t)3 * N * (C + 8) * 2 + 8 * its::AS_QB * 64 * 4;
  if (N <= 32 || smem_need > 227 * 1024) {      // tiny maps, or q|k|v of one image do not fit shared memory (C = 1024)
    ITS_REQUIRE(n_img <= 65535, "its_attention_small: n_img");
    ITS_LAUNCH(its::attention_tiny_kernel, dim3(N, n_img), dim3(256), 0, its::as_stream(stream),
               static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(qkv), N, C, scale);
    ITS_CHECK_LAUNCH();
    return ITS_OK;
  }
  const size_t smem = (size_t)3 * N * (C + 8) * 2 + 8 * its::AS_QB * 64 * 4;
  ITS_REQUIRE(smem <= 227 * 1024, "its_attention_small: N=%d C=%d needs %zu bytes of shared memory", N, C, smem);
  static its::PerDeviceBytes configured;
  if (configured.need(smem))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(its::attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ITS_LAUNCH(its::attention_small_kernel, dim3(n_img), dim3(256), smem, its::as_stream(stream),
      static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(qkv), N, C, scale);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}
