// Verifier statistics, per-candidate scores and first-index argmax
// (search/verifier.py:45-66, 207-248, 262-287; search/search_algorithm.py:79-81).
// All reductions are warp-shuffle + one shared-memory hop.
#include "its_common.cuh"

namespace its {

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? s_red[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread of every warp holds the total
}
__device__ __forceinline__ float block_min(float v, float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? s_red[lane] : INFINITY;
  t = warp_min(t);
  return t;
}

// One CTA per image.  stats[img] = {mean, unbiased var, min, |pooled|_2 (0 without feats)};
// feats[img][C*64] = L2-normalised adaptive 8x8 average pool (may be NULL).
__global__ void __launch_bounds__(256) image_stats_kernel(float* __restrict__ stats,
                                                          float* __restrict__ feats,
                                                          const float* __restrict__ images, int C,
                                                          int H, int W) {
  pdl_prologue();
  __shared__ float s_red[32];
  __shared__ float s_feat[4 * 64];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int n = C * H * W;
  const float* x = images + (long long)img * n;
  float s = 0.f, mn = INFINITY;
  for (int i = tid; i < n; i += blockDim.x) {
    const float v = x[i];
    s += v;
    mn = fminf(mn, v);
  }
  const float mean = block_sum(s, s_red) / (float)n;
  mn = block_min(mn, s_red);
  float m2 = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    const float d = x[i] - mean;
    m2 = fmaf(d, d, m2);
  }
  m2 = block_sum(m2, s_red);
  if (tid == 0) {
    float4 o = make_float4(mean, m2 / (float)(n - 1), mn, 0.f);
    reinterpret_cast<float4*>(stats)[img] = o;
  }
  if (feats == nullptr) return;
  // adaptive_avg_pool2d -> (8, 8): window [floor(i*H/8), ceil((i+1)*H/8))
  const int nf = C * 64;
  float sq = 0.f;
  for (int f = tid; f < nf; f += blockDim.x) {
    const int c = f >> 6, i = (f >> 3) & 7, j = f & 7;
    const int y0 = (i * H) / 8, y1 = ((i + 1) * H + 7) / 8;
    const int x0 = (j * W) / 8, x1 = ((j + 1) * W + 7) / 8;
    float a = 0.f;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) a += x[((long long)c * H + yy) * W + xx];
    a /= (float)((y1 - y0) * (x1 - x0));
    s_feat[f] = a;
    sq = fmaf(a, a, sq);
  }
  sq = block_sum(sq, s_red);
  if (tid == 0) stats[(long long)img * 4 + 3] = sqrtf(sq);  // norm of the pooled vector
  const float inv = 1.0f / fmaxf(sqrtf(sq), 1e-12f);         // F.normalize eps
  for (int f = tid; f < nf; f += blockDim.x) feats[(long long)img * nf + f] = s_feat[f] * inv;
}

// One warp per candidate.
__global__ void __launch_bounds__(128) candidate_scores_kernel(float* __restrict__ scores,
                                                               const float* __restrict__ stats,
                                                               const float* __restrict__ feats,
                                                               int n_cand, int per_cand,
                                                               int feat_dim, int kind) {
  pdl_prologue();
  const int cand = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cand >= n_cand) return;
  const float4* st = reinterpret_cast<const float4*>(stats) + (long long)cand * per_cand;
  float score;
  if (kind == 0) {
    float v = 0.f;
    for (int b = lane; b < per_cand; b += 32) v += st[b].y;
    v = warp_sum(v) / (float)per_cand;
    score = 1.0f / (1.0f + v);
  } else if (kind == 1) {
    float mn = INFINITY, sd = 0.f;
    for (int b = lane; b < per_cand; b += 32) {
      mn = fminf(mn, st[b].z);
      sd += sqrtf(st[b].y);
    }
    mn = warp_min(mn);
    sd = warp_sum(sd) / (float)per_cand;
    if (mn < 0.f) sd *= 0.5f;  // std((x+1)/2) = std(x)/2
    score = sd + sd;           // color_diversity + contrast are the same quantity
  } else {
    const float* f = feats + (long long)cand * per_cand * feat_dim;
    float acc = 0.f;
    const int pairs = per_cand * per_cand;
    for (int pr = 0; pr < pairs; ++pr) {
      const int i = pr / per_cand, j = pr - i * per_cand;
      if (i == j) continue;
      float d = 0.f;
      for (int e = lane; e < feat_dim; e += 32) d = fmaf(f[(long long)i * feat_dim + e], f[(long long)j * feat_dim + e], d);
      acc += warp_sum(d);
    }
    const int cnt = pairs - per_cand;
    score = (cnt > 0) ? acc / (float)cnt : __int_as_float(0x7fc00000);  // mean of empty = NaN
  }
  if (lane == 0) scores[cand] = score;
}

struct Best { float v; int i; };
__device__ __forceinline__ Best better(Best a, Best b) {
  // strict '>' with first index on ties; i < 0 means "none yet"
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__global__ void __launch_bounds__(1024) argmax_first_kernel(int* __restrict__ idx_out,
                                                            float* __restrict__ val_out,
                                                            const float* __restrict__ scores, int n) {
  pdl_prologue();
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  Best b = {-INFINITY, -1};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = scores[i];
    if (v > b.v) { b.v = v; b.i = i; }  // NaN and -inf never win, ties keep the earlier index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best t = {__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
    b = better(b, t);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_v[warp] = b.v; s_i[warp] = b.i; }
  __syncthreads();
  if (warp == 0) {
    Best t = {-INFINITY, -1};
    if (lane < (int)(blockDim.x >> 5)) { t.v = s_v[lane]; t.i = s_i[lane]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best u = {__shfl_xor_sync(0xffffffffu, t.v, o), __shfl_xor_sync(0xffffffffu, t.i, o)};
      t = better(t, u);
    }
    if (lane == 0) { *idx_out = t.i; *val_out = t.v; }
  }
}

// Top-k selection, same ordering rule applied k times: rank r is the first index of the largest score that comes
// after rank r-1 in the order (score descending, index ascending).  NaN and -inf never rank; ranks beyond the
// number of eligible scores get index -1.  One block, warp-shuffle + shared-memory reduction per rank.
__global__ void __launch_bounds__(1024) topk_first_kernel(int* __restrict__ idx_out, float* __restrict__ val_out,
                                                          const float* __restrict__ scores, int n, int k) {
  pdl_prologue();
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  __shared__ Best s_prev;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Best prev = {INFINITY, -1};
  for (int r = 0; r < k; ++r) {
    Best b = {-INFINITY, -1};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = scores[i];
      const bool after = r == 0 || v < prev.v || (v == prev.v && i > prev.i);
      if (after && v > b.v) { b.v = v; b.i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best t = {__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
      b = better(b, t);
    }
    if (lane == 0) { s_v[warp] = b.v; s_i[warp] = b.i; }
    __syncthreads();
    if (warp == 0) {
      Best t = {-INFINITY, -1};
      if (lane < (int)(blockDim.x >> 5)) { t.v = s_v[lane]; t.i = s_i[lane]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Best u = {__shfl_xor_sync(0xffffffffu, t.v, o), __shfl_xor_sync(0xffffffffu, t.i, o)};
        t = better(t, u);
      }
      if (lane == 0) {
        idx_out[r] = t.i;
        val_out[r] = t.v;
        s_prev = t;
      }
    }
    __syncthreads();
    prev = s_prev;
    if (prev.i < 0) {            // nothing left: the remaining ranks are empty
      for (int q = r + 1 + threadIdx.x; q < k; q += blockDim.x) { idx_out[q] = -1; val_out[q] = -INFINITY; }
      break;
    }
  }
}

}  // namespace its

extern "C" int its_image_stats(float* stats, float* feats, const float* images, int32_t n_img,
                               int32_t C, int32_t H, int32_t W, void* stream) {
  ITS_REQUIRE(stats && images, "its_image_stats: null pointer");
  ITS_REQUIRE(n_img > 0 && C > 0 && C <= 4 && H > 0 && W > 0 && (long long)C * H * W > 1,
              "its_image_stats: unsupported shape C=%d H=%d W=%d", C, H, W);
  ITS_LAUNCH(its::image_stats_kernel, dim3(n_img), dim3(256), 0, its::as_stream(stream), stats, feats, images, C, H, W);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_candidate_scores(float* scores, const float* stats, const float* feats,
                                    int32_t n_cand, int32_t per_cand, int32_t feat_dim,
                                    int32_t kind, void* stream) {
  ITS_REQUIRE(scores && stats, "its_candidate_scores: null pointer");
  ITS_REQUIRE(n_cand > 0 && per_cand > 0 && kind >= 0 && kind <= 2, "its_candidate_scores: bad arguments");
  ITS_REQUIRE(kind != 2 || (feats != nullptr && feat_dim > 0), "its_candidate_scores: kind 2 needs feats");
  const int blocks = (n_cand * 32 + 127) / 128;
  ITS_LAUNCH(its::candidate_scores_kernel, dim3(blocks), dim3(128), 0, its::as_stream(stream), scores, stats, feats, n_cand,
                                                                          per_cand, feat_dim, kind);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_argmax_first(int32_t* idx_out, float* val_out, const float* scores, int32_t n,
                                void* stream) {
  ITS_REQUIRE(idx_out && val_out && scores && n > 0, "its_argmax_first: bad arguments");
  ITS_LAUNCH(its::argmax_first_kernel, dim3(1), dim3(1024), 0, its::as_stream(stream), idx_out, val_out, scores, n);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_topk_first(int32_t* idx_out, float* val_out, const float* scores, int32_t n, int32_t k,
                              void* stream) {
  ITS_REQUIRE(idx_out && val_out && scores && n > 0 && k > 0, "its_topk_first: bad arguments");
  ITS_LAUNCH(its::topk_first_kernel, dim3(1), dim3(1024), 0, its::as_stream(stream), idx_out, val_out, scores, n, k);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}
