// fp32-grade path ("precision = fp32"): the reference runs everything in fp32 (Diffusion/Diffusion.py:74-99,
// Model.py, ModelCondition.py); north_star asks for samples within 1e-4 of it in that mode.  These are plain
// CUDA-core kernels over NHWC fp32 tensors, one per operation of the reference's UNet, with fp32 FMAs in a
// fixed order and double-precision GroupNorm statistics.  They are the parity instrument (and an independent
// on-device cross-check of the tcgen05 path), not the throughput path: nothing here touches the tensor cores.
#include "its_common.cuh"

namespace its {

// out[b][y][x][c] = in[b % n_img_in][c][y][x]   (the sampler state is NCHW; CFG: both halves read the same x_t)
__global__ void __launch_bounds__(256) f32_nchw_to_nhwc_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                                long long total, int n_img_in, int C, int H, int W) {
  pdl_prologue();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    out[i] = in[(((b % n_img_in) * C + c) * H + y) * W + x];
  }
}

struct F32ConvArgs {
  float* out;
  const float* in0;
  const float* in1;
  const float* Wt;
  const float* bias;
  const float* vec;
  const float* vec2;
  const float* res;
  int C0, C1, B, Hin, Win, Hout, Wout, Cout, CoutP, k, stride, mode, out_nchw, vec_stride, vec2_stride;
};

// Direct convolution over the channel concatenation of up to two NHWC tensors.
//   mode 0: nn.Conv2d(k, stride, padding k/2)                                     (Model.py:99,173,190,193)
//   mode 1: F.interpolate(scale 2, nearest) then nn.Conv2d(k, 1, k/2)              (Model.py:122-125)
//   mode 2: nn.ConvTranspose2d(k, stride 2, padding k/2, output_padding 1)        (ModelCondition.py:80)
// Wt is [k*k][C0+C1][CoutP] (CoutP = Cout rounded up to 4): a thread owns one output pixel and four consecutive
// output channels; a warp reads one activation (broadcast) and 32 float4 weights (coalesced) per FMA group.
__global__ void __launch_bounds__(256) f32_conv2d_kernel(const F32ConvArgs a) {
  pdl_prologue();
  const int n4 = a.CoutP >> 2;
  const int Cin = a.C0 + a.C1;
  const int pad = a.k >> 1;
  const long long total = (long long)a.B * a.Hout * a.Wout * n4;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += gridDim.x * 256LL) {
    const int nq = (int)(idx % n4);
    const long long pix = idx / n4;
    const int x = (int)(pix % a.Wout), y = (int)((pix / a.Wout) % a.Hout);
    const long long b = pix / ((long long)a.Wout * a.Hout);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int ky = 0; ky < a.k; ++ky) {
      int iy;
      bool oky;
      if (a.mode == 0) {
        iy = y * a.stride + ky - pad;
        oky = iy >= 0 && iy < a.Hin;
      } else if (a.mode == 1) {
        const int vy = y + ky - pad;
        oky = vy >= 0 && vy < 2 * a.Hin;
        iy = vy >> 1;
      } else {
        const int ty = y + pad - ky;
        oky = ty >= 0 && (ty & 1) == 0 && (ty >> 1) < a.Hin;
        iy = ty >> 1;
      }
      if (!oky) continue;
      for (int kx = 0; kx < a.k; ++kx) {
        int ix;
        bool okx;
        if (a.mode == 0) {
          ix = x * a.stride + kx - pad;
          okx = ix >= 0 && ix < a.Win;
        } else if (a.mode == 1) {
          const int vx = x + kx - pad;
          okx = vx >= 0 && vx < 2 * a.Win;
          ix = vx >> 1;
        } else {
          const int tx = x + pad - kx;
          okx = tx >= 0 && (tx & 1) == 0 && (tx >> 1) < a.Win;
          ix = tx >> 1;
        }
        if (!okx) continue;
        const float* w = a.Wt + ((long long)(ky * a.k + kx) * Cin) * a.CoutP + nq * 4;
        const long long ipix = (b * a.Hin + iy) * a.Win + ix;
        // channels in ascending order, one FMA chain per output channel (the summation order is the same
        // whether a source is read four channels at a time or one by one)
        auto run = [&](const float* src, int C, const float* wsrc) {
          int c = 0;
          if ((C & 3) == 0) {
            for (; c < C; c += 4) {
              const float4 v4 = __ldg(reinterpret_cast<const float4*>(src + c));
              const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wsrc + (long long)(c + j) * a.CoutP));
                acc0 = fmaf(vv[j], w4.x, acc0); acc1 = fmaf(vv[j], w4.y, acc1);
                acc2 = fmaf(vv[j], w4.z, acc2); acc3 = fmaf(vv[j], w4.w, acc3);
              }
            }
          }
          for (; c < C; ++c) {
            const float v = __ldg(src + c);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wsrc + (long long)c * a.CoutP));
            acc0 = fmaf(v, w4.x, acc0); acc1 = fmaf(v, w4.y, acc1); acc2 = fmaf(v, w4.z, acc2); acc3 = fmaf(v, w4.w, acc3);
          }
        };
        run(a.in0 + ipix * a.C0, a.C0, w);
        if (a.C1 > 0) run(a.in1 + ipix * a.C1, a.C1, w + (long long)a.C0 * a.CoutP);
      }
    }
    const float acc[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nq * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[j];
      if (a.bias) v += a.bias[n];
      if (a.vec) v += a.vec[b * a.vec_stride + n];
      if (a.vec2) v += a.vec2[b * a.vec2_stride + n];
      if (a.res) v += a.res[pix * a.Cout + n];
      if (a.out_nchw)
        a.out[((b * a.Cout + n) * a.Hout + y) * a.Wout + x] = v;
      else
        a.out[pix * a.Cout + n] = v;
    }
  }
}

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];   // fixed order
  return t;
}

// nn.GroupNorm(groups, C0+C1) (+ Swish) over the channel concatenation of up to two NHWC fp32 tensors
// (Model.py:132,170-173,186-190,257-259 after the concat of Model.py:279-280).  One CTA per (group, image):
// mean, then the centred sum of squares, both accumulated in double; biased variance like nn.GroupNorm.
__global__ void __launch_bounds__(256) f32_group_norm_kernel(float* __restrict__ out, const float* __restrict__ in0, int C0,
                                                              const float* __restrict__ in1, int C1,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int HW, int groups,
                                                              float eps, int silu) {
  __shared__ double sh[8];
  pdl_prologue();
  const int g = blockIdx.x;
  const long long b = blockIdx.y;
  const int C = C0 + C1, cg = C / groups, c0 = g * cg;
  const int n = cg * HW;
  auto load = [&](int e) -> float {
    const int p = e / cg, c = c0 + (e - p * cg);
    return c < C0 ? in0[(b * HW + p) * C0 + c] : in1[(b * HW + p) * C1 + (c - C0)];
  };
  double s = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) s += (double)load(e);
  const double mean = block_sum_double(s, sh) / (double)n;
  double q = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const double d = (double)load(e) - mean;
    q += d * d;
  }
  const double var = block_sum_double(q, sh) / (double)n;
  const float fmean = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int p = e / cg, c = c0 + (e - p * cg);
    float y = (load(e) - fmean) * rstd * gamma[c] + beta[c];
    if (silu) y = y / (1.0f + expf(-y));
    out[(b * HW + p) * C + c] = y;
  }
}

// Single-head attention core (Model.py:147-161): out[b, i, :] = softmax_j(scale * q_i . k_j) v_j.  qkv is the fused
// projection tensor [n_img][N][3C] (q | k | v).  One warp per (image, query): scores in shared memory, exp / sum
// in fp32 with expf, probabilities normalised before the product with V (as F.softmax then torch.bmm do).
__global__ void __launch_bounds__(256) f32_attention_kernel(float* __restrict__ out, const float* __restrict__ qkv, int N,
                                                             int C, float scale) {
  extern __shared__ float sm[];
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  const long long b = blockIdx.y;
  if (i >= N) return;
  float* q = sm + (long long)warp * (C + N);
  float* s = q + C;
  const float* base = qkv + b * N * 3 * C;
  for (int c = lane; c < C; c += 32) q[c] = base[(long long)i * 3 * C + c];
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    const float* kr = base + (long long)j * 3 * C + C;
    float d = 0.f;
    for (int c = 0; c < C; ++c) d = fmaf(q[c], kr[c], d);
    d *= scale;
    s[j] = d;
    mx = fmaxf(mx, d);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float tot = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    tot += e;
  }
  tot = warp_sum(tot);
  __syncwarp();
  const float inv = 1.0f / tot;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(s[j] * inv, base[(long long)j * 3 * C + 2 * C + c], acc);
    out[(b * N + i) * C + c] = acc;
  }
}

}  // namespace its

extern "C" int its_f32_nchw_to_nhwc(float* out, const float* in, int32_t n_img, int32_t n_img_in, int32_t C, int32_t H,
                                    int32_t W, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && in && n_img > 0 && n_img_in > 0 && C > 0 && H > 0 && W > 0, "its_f32_nchw_to_nhwc: bad arguments");
  const long long total = (long long)n_img * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ITS_LAUNCH(f32_nchw_to_nhwc_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), out, in, total, n_img_in,
             C, H, W);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_f32_conv2d(float* out, const float* in0, int32_t C0, const float* in1, int32_t C1, const float* Wt,
                              const float* bias, const float* vec, int32_t vec_stride, const float* vec2,
                              int32_t vec2_stride, const float* res, int32_t B, int32_t Hin, int32_t Win, int32_t Cout,
                              int32_t k, int32_t stride, int32_t mode, int32_t out_nchw, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && in0 && Wt, "its_f32_conv2d: null pointer");
  ITS_REQUIRE(C0 > 0 && C1 >= 0 && (C1 == 0 || in1 != nullptr), "its_f32_conv2d: channels %d + %d", C0, C1);
  ITS_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Cout > 0, "its_f32_conv2d: bad shape");
  ITS_REQUIRE((k == 1 || k == 3 || k == 5) && (stride == 1 || stride == 2), "its_f32_conv2d: k=%d stride=%d", k, stride);
  ITS_REQUIRE(mode >= 0 && mode <= 2 && (mode == 0 || stride == (mode == 2 ? 2 : 1)),
              "its_f32_conv2d: mode=%d with stride=%d", mode, stride);
  ITS_REQUIRE(!(out_nchw && res), "its_f32_conv2d: a residual needs the NHWC output layout");
  ITS_REQUIRE((reinterpret_cast<uintptr_t>(Wt) & 15) == 0, "its_f32_conv2d: weight alignment");
  F32ConvArgs a;
  a.out = out; a.in0 = in0; a.in1 = in1; a.Wt = Wt; a.bias = bias; a.vec = vec; a.vec2 = vec2; a.res = res;
  a.C0 = C0; a.C1 = C1; a.B = B; a.Hin = Hin; a.Win = Win; a.Cout = Cout; a.CoutP = (Cout + 3) & ~3;
  a.k = k; a.stride = stride; a.mode = mode; a.out_nchw = out_nchw; a.vec_stride = vec_stride; a.vec2_stride = vec2_stride;
  if (mode == 0) {
    ITS_REQUIRE(Hin % stride == 0 && Win % stride == 0, "its_f32_conv2d: odd map for stride 2");
    a.Hout = Hin / stride; a.Wout = Win / stride;
  } else {
    a.Hout = 2 * Hin; a.Wout = 2 * Win;
  }
  const long long total = (long long)B * a.Hout * a.Wout * (a.CoutP >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ITS_LAUNCH(f32_conv2d_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), a);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_f32_group_norm(float* out, const float* in0, int32_t C0, const float* in1, int32_t C1,
                                  const float* gamma, const float* beta, int32_t n_img, int32_t HW, int32_t groups,
                                  float eps, int32_t silu, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && in0 && gamma && beta && (C1 == 0 || in1), "its_f32_group_norm: null pointer");
  ITS_REQUIRE(C0 > 0 && C1 >= 0 && groups > 0 && (C0 + C1) % groups == 0 && n_img > 0 && HW > 0,
              "its_f32_group_norm: C=%d+%d groups=%d", C0, C1, groups);
  ITS_LAUNCH(f32_group_norm_kernel, dim3((unsigned)groups, (unsigned)n_img), dim3(256), 0, as_stream(stream), out, in0,
             C0, in1, C1, gamma, beta, HW, groups, eps, silu);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_f32_attention(float* out, const float* qkv, int32_t n_img, int32_t N, int32_t C, float scale,
                                 void* stream) {
  using namespace its;
  ITS_REQUIRE(out && qkv && n_img > 0 && N > 0 && C > 0, "its_f32_attention: bad arguments");
  const size_t smem = (size_t)8 * (C + N) * sizeof(float);
  ITS_REQUIRE(smem <= 200 * 1024, "its_f32_attention: N=%d C=%d needs %zu bytes of shared memory", N, C, smem);
  static PerDeviceBytes configured;
  if (smem > 48 * 1024 && configured.need(smem))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(f32_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ITS_LAUNCH(f32_attention_kernel, dim3((unsigned)((N + 7) / 8), (unsigned)n_img), dim3(256), smem, as_stream(stream), out,
             qkv, N, C, scale);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}
