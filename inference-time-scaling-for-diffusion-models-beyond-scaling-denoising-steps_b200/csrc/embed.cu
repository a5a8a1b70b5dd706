// Time / label embeddings and the small fp32 linears around them.
#include "its_common.cuh"

namespace its {

// emb[b, 2i] = sin(t * f_i), emb[b, 2i+1] = cos(t * f_i)   (Model.py:76-88)
__global__ void time_embed_kernel(float* __restrict__ out, const long long* __restrict__ t_idx,
                                  const int* __restrict__ t_dev, const float* __restrict__ freq,
                                  int n_rows, int half) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * half) return;
  const int b = i / half, j = i - b * half;
  const float t = (t_idx != nullptr) ? (float)t_idx[b] : (float)(*t_dev);
  const float a = t * freq[j];
  // accurate sinf/cosf: arguments reach T (thousands of radians)
  out[(long long)b * 2 * half + 2 * j] = sinf(a);
  out[(long long)b * 2 * half + 2 * j + 1] = cosf(a);
}

__global__ void embed_rows_kernel(float* __restrict__ out, const float* __restrict__ table,
                                  const long long* __restrict__ idx, const int* __restrict__ t_dev,
                                  int n_rows, int dim, int n_table_rows) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * dim) return;
  const int b = i / dim, j = i - b * dim;
  const long long r = (idx != nullptr) ? idx[b] : (long long)(*t_dev);
  // nn.Embedding raises on an index outside the table (the host shells check it); the kernel never reads outside
  // the table and marks such a row with NaN, which the sampler's "nan in tensor." assertion reports
  out[i] = (r < 0 || r >= n_table_rows) ? __int_as_float(0x7fc00000) : table[r * (long long)dim + j];
}

// y[b, n] = sum_k act(x[b,k]) W[n,k] + bias[n]; one warp per (b, n): lanes stride
// over K (coalesced W reads), shuffle reduce.  Tiny matrices only.
__global__ void __launch_bounds__(256) linear_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                     const float* __restrict__ W,
                                                     const float* __restrict__ bias, int n_rows,
                                                     int K, int N, int silu_in, int silu_out,
                                                     int accumulate) {
  pdl_prologue();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)n_rows * N) return;
  const int b = (int)(warp / N), n = (int)(warp - (long long)b * N);
  const float* xr = x + (long long)b * K;
  const float* wr = W + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = xr[k];
    if (silu_in) xv = xv / (1.0f + expf(-xv));
    acc = fmaf(xv, __ldg(wr + k), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (bias ? bias[n] : 0.f);
    if (silu_out) v = v / (1.0f + expf(-v));
    float* dst = y + (long long)b * N + n;
    *dst = accumulate ? (*dst + v) : v;
  }
}

// Same sums for R rows at a time: the activated input rows are staged once per CTA in shared
// memory and every weight row is read once per R rows instead of once per row (the label
// projections of the conditional net have one row per image).  Per (b, n) the summation order is
// the one of linear_kernel (lane-strided over K, then the shuffle tree), so results are bit-identical.
template <int R>
__global__ void __launch_bounds__(256) linear_rows_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                          const float* __restrict__ W,
                                                          const float* __restrict__ bias, int n_rows, int K,
                                                          int N, int silu_in, int silu_out, int accumulate,
                                                          int cols_per_block) {
  extern __shared__ float xs[];   // [R][K]
  pdl_prologue();
  const int b0 = blockIdx.y * R;
  for (int i = threadIdx.x; i < R * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    float xv = (b0 + r < n_rows) ? x[(long long)(b0 + r) * K + k] : 0.f;
    if (silu_in) xv = xv / (1.0f + expf(-xv));
    xs[i] = xv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_end = min(N, (int)(blockIdx.x + 1) * cols_per_block);
  for (int n = blockIdx.x * cols_per_block + warp; n < n_end; n += 8) {
    const float* wr = W + (long long)n * K;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(xs[r * K + k], wv, acc[r]);
    }
    const float bn = bias ? bias[n] : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float a = warp_sum(acc[r]);
      if (lane == 0 && b0 + r < n_rows) {
        float v = a + bn;
        if (silu_out) v = v / (1.0f + expf(-v));
        float* dst = y + (long long)(b0 + r) * N + n;
        *dst = accumulate ? (*dst + v) : v;
      }
    }
  }
}

}  // namespace its

extern "C" int its_time_embed(float* out, const int64_t* t_idx, const int32_t* t_dev,
                              const float* freq, int32_t n_rows, int32_t d_model, void* stream) {
  ITS_REQUIRE(out && freq && (t_idx || t_dev), "its_time_embed: null pointer");
  ITS_REQUIRE(n_rows > 0 && d_model > 0 && d_model % 2 == 0, "its_time_embed: bad d_model=%d", d_model);
  const int half = d_model / 2, total = n_rows * half;
  ITS_LAUNCH(its::time_embed_kernel, dim3((total + 127) / 128), dim3(128), 0, its::as_stream(stream), 
      out, reinterpret_cast<const long long*>(t_idx), t_dev, freq, n_rows, half);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_embed_rows(float* out, const float* table, const int64_t* idx, const int32_t* t_dev,
                              int32_t n_rows, int32_t dim, int32_t n_table_rows, void* stream) {
  ITS_REQUIRE(out && table && (idx || t_dev), "its_embed_rows: null pointer");
  ITS_REQUIRE(n_rows > 0 && dim > 0 && n_table_rows > 0, "its_embed_rows: bad sizes");
  const int total = n_rows * dim;
  ITS_LAUNCH(its::embed_rows_kernel, dim3((total + 127) / 128), dim3(128), 0, its::as_stream(stream), 
      out, table, reinterpret_cast<const long long*>(idx), t_dev, n_rows, dim, n_table_rows);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_linear(float* y, const float* x, const float* W, const float* bias,
                          int32_t n_rows, int32_t K, int32_t N, int32_t silu_in, int32_t silu_out,
                          int32_t accumulate, void* stream) {
  ITS_REQUIRE(y && x && W, "its_linear: null pointer");
  ITS_REQUIRE(n_rows > 0 && K > 0 && N > 0, "its_linear: bad sizes");
  if (n_rows > 1 && (long long)K * 16 * 4 <= 48 * 1024) {
    // row-blocked variant: 64 output columns per CTA
    const int cpb = 64;
    if (n_rows > 8) {
      ITS_LAUNCH(its::linear_rows_kernel<16>, dim3((N + cpb - 1) / cpb, (n_rows + 15) / 16), dim3(256),
                 (size_t)K * 16 * 4, its::as_stream(stream), y, x, W, bias, n_rows, K, N, silu_in, silu_out,
                 accumulate, cpb);
    } else {
      ITS_LAUNCH(its::linear_rows_kernel<4>, dim3((N + cpb - 1) / cpb, (n_rows + 3) / 4), dim3(256),
                 (size_t)K * 4 * 4, its::as_stream(stream), y, x, W, bias, n_rows, K, N, silu_in, silu_out,
                 accumulate, cpb);
    }
    ITS_CHECK_LAUNCH();
    return ITS_OK;
  }
  const long long warps = (long long)n_rows * N;
  const long long blocks = (warps * 32 + 255) / 256;
  ITS_REQUIRE(blocks < (1LL << 31), "its_linear: too large");
  ITS_LAUNCH(its::linear_kernel, dim3((unsigned)blocks), dim3(256), 0, its::as_stream(stream), 
      y, x, W, bias, n_rows, K, N, silu_in, silu_out, accumulate);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}
