// CUDA-core convolutions: the Cin=3 head and Cout=3 tail of the UNet (too thin
// for a 128-wide tensor-core tile) and a plain reference implementation of the
// tap-GEMM the tcgen05 kernel computes (debug twin + channel counts that are not
// multiples of 64).
#include "tapgemm.cuh"
#include <cuda_fp16.h>

namespace its {

// ---------------------------------------------------------------- head ----
// out NHWC bf16 [n_img][H][W][Cout]; x NCHW fp32 [n_img_in][Cin][H][W];
// W OIHW fp32 [Cout][Cin][3][3].  One thread = one pixel: its Cin*9 inputs stay in
// registers, the weights are read from shared memory as warp-wide broadcasts
// (a warp = 32 consecutive pixels, all on the same 16 output channels).
template <int CIN>
__global__ void __launch_bounds__(256) conv_head_kernel(
    __nv_bfloat16* __restrict__ out, const float* __restrict__ x, const float* __restrict__ W,
    const float* __restrict__ bias, int n_img, int n_img_in, int H, int Wd, int Cout) {
  pdl_prologue();
  extern __shared__ float s_w[];  // [CIN*9][Cout] then bias[Cout]
  constexpr int K = CIN * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int co = i / K, k = i - co * K;  // W[co][ci][ky][kx] flat = co*K + k
    s_w[k * Cout + co] = W[i];
  }
  float* s_b = s_w + K * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long npix = (long long)n_img * H * Wd;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(pix % Wd);
    const int y = (int)((pix / Wd) % H);
    const int b = (int)(pix / ((long long)Wd * H));
    const float* xin = x + (long long)(b % n_img_in) * CIN * H * Wd;
    float in[K];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = y + ky - 1, xx = xw + kx - 1;
          const bool inb = (yy >= 0 && yy < H && xx >= 0 && xx < Wd);
          in[(ci * 3 + ky) * 3 + kx] = inb ? __ldg(xin + ((long long)ci * H + yy) * Wd + xx) : 0.f;
        }
    __nv_bfloat16* orow = out + pix * Cout;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = s_b[c0 + j];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + k * Cout + c0);
        const float v = in[k];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w4 = wr[j4];
          acc[4 * j4 + 0] = fmaf(v, w4.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(v, w4.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(v, w4.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(v, w4.w, acc[4 * j4 + 3]);
        }
      }
      *reinterpret_cast<bf16x8*>(orow + c0) = pack8(acc);
      *reinterpret_cast<bf16x8*>(orow + c0 + 8) = pack8(acc + 8);
    }
  }
}


// ------------------------------------------------------- head patches ----
// The Cin=3 head as a tensor-core GEMM: every output pixel's 3x3xCin input patch is written as
// one 128-channel bf16 "pixel" [x_hi | x_lo | x_hi | 0...] (x_hi = bf16(x), x_lo = bf16(x - x_hi)),
// to be multiplied with the packed weights [w_hi | w_hi | w_lo | 0] by a 1x1 tap-GEMM: the three
// products x_hi*w_hi + x_lo*w_hi + x_hi*w_lo keep ~16 mantissa bits of the fp32 convolution
// (Model.py:269 runs it in fp32; x_t reaches +-30 at early steps).
// One thread gathers the 9*Cin values of one pixel's patch (zero outside the image), splits them
// into hi / lo and writes the 256-byte patch row into a per-warp staging tile; the warp then streams
// its 32 consecutive patch rows (8 KB, contiguous in global memory) out with coalesced 16-byte stores.
constexpr int HP_THREADS = 128;
template <int Cin>
__global__ void __launch_bounds__(HP_THREADS) head_patches_kernel(__nv_bfloat16* __restrict__ out,
                                                                  const float* __restrict__ x, long long npix,
                                                                  int n_img_in, int H, int Wd) {
  __shared__ uint4 s_tile[HP_THREADS / 32][32][17];      // pitch 17: conflict-free both ways
  pdl_prologue();
  constexpr int K = Cin * 9;                             // 3K <= 128
  const int HW = H * Wd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long base = (blockIdx.x * (long long)(HP_THREADS / 32) + warp) * 32; base < npix;
       base += (long long)gridDim.x * HP_THREADS) {
    const long long pix = base + lane;
    __nv_bfloat16 row[128];
#pragma unroll
    for (int k = 0; k < 128; ++k) row[k] = __float2bfloat16_rn(0.f);
    if (pix < npix) {
      const int b = (int)(pix / HW);
      const int rem = (int)(pix - (long long)b * HW);
      const int y = rem / Wd, xw = rem - y * Wd;
      const float* xin = x + (long long)(b % n_img_in) * Cin * HW;
#pragma unroll
      for (int t = 0; t < K; ++t) {
        {
          const int ci = t / 9, r = t - ci * 9, ky = r / 3, kx = r - ky * 3;
          const int yy = y + ky - 1, xx = xw + kx - 1;
          float xv = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < Wd) xv = __ldg(xin + ci * HW + yy * Wd + xx);
          const __nv_bfloat16 hi = __float2bfloat16_rn(xv);
          const __nv_bfloat16 lo = __float2bfloat16_rn(xv - __bfloat162float(hi));
          row[t] = hi;
          if (K + t < 128) row[K + t] = lo;
          if (2 * K + t < 128) row[2 * K + t] = hi;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < 16; ++v) s_tile[warp][lane][v] = *reinterpret_cast<const uint4*>(&row[v * 8]);
    __syncwarp();
    // 32 rows x 16 vectors = 512 consecutive 16-byte words of global memory
    uint4* dst = reinterpret_cast<uint4*>(out) + base * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int w = j * 32 + lane;
      if (base + (w >> 4) < npix) dst[w] = s_tile[warp][w >> 4][w & 15];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- tail ----
// act NHWC bf16 [n_img][H][W][Cin] (GroupNorm+Swish already applied);
// out NCHW fp32 [n_img][Cout][H][W], Cout <= 4.  One warp per pixel.
__global__ void __launch_bounds__(256) conv_tail_kernel(
    float* __restrict__ out, const __nv_bfloat16* __restrict__ act, const float* __restrict__ W,
    const float* __restrict__ bias, int n_img, int H, int Wd, int Cin, int Cout) {
  pdl_prologue();
  extern __shared__ float s_w[];  // [9][Cin][4]
  for (int i = threadIdx.x; i < 9 * Cin * 4; i += blockDim.x) {
    const int co = i & 3, ci = (i >> 2) % Cin, tap = (i >> 2) / Cin;
    s_w[i] = (co < Cout) ? W[((long long)co * Cin + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = Cin / 8;
  const long long npix = (long long)n_img * H * Wd;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < npix;
       pix += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int xw = (int)(pix % Wd);
    const int y = (int)((pix / Wd) % H);
    const long long b = pix / ((long long)Wd * H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int e = lane; e < 9 * nv; e += 32) {
      const int tap = e / nv, cv = e - tap * nv;
      const int yy = y + tap / 3 - 1, xx = xw + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= Wd) continue;
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(act + ((b * H + yy) * Wd + xx) * Cin + cv * 8), f);
      const float4* wr = reinterpret_cast<const float4*>(s_w + ((long long)tap * Cin + cv * 8) * 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 wv = wr[j];
        a0 = fmaf(f[j], wv.x, a0);
        a1 = fmaf(f[j], wv.y, a1);
        a2 = fmaf(f[j], wv.z, a2);
        a3 = fmaf(f[j], wv.w, a3);
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane < Cout) {
      const float v = (lane == 0 ? a0 : lane == 1 ? a1 : lane == 2 ? a2 : a3) + (bias ? bias[lane] : 0.f);
      out[((b * Cout + lane) * H + y) * Wd + xw] = v;
    }
  }
}

// ------------------------------------------------- reference tap-GEMM ----
// One thread per (GEMM row, output column); same parameter block, packed
// weights and epilogue as the tcgen05 kernel.
__global__ void __launch_bounds__(128) tapgemm_ref_kernel(const TapGemmParams p) {
  pdl_prologue();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  const long long row = blockIdx.x;
  const DevPhase& ph = p.phase[blockIdx.z];
  if (n >= p.Cout) return;
  const int x = (int)(row % p.Wm);
  const int y = (int)((row / p.Wm) % p.Hm);
  const int b = (int)(row / ((long long)p.Wm * p.Hm));
  const __nv_bfloat16* wrow = p.w + (long long)b * p.w_batch_stride + (long long)n * p.w_pitch + ph.w_k0;
  float acc = 0.f;
  int k = 0;
  for (int t = 0; t < ph.ntaps; ++t) {
    const DevSrc& s = p.src[ph.src[t]];
    const int yy = y * s.stride + ph.dy[t], xx = x * s.stride + ph.dx[t];
    const bool inb = (yy >= 0 && yy < s.H && xx >= 0 && xx < s.W);
    if (inb) {
      const __nv_bfloat16* a =
          s.ptr + (((long long)(s.bcast ? 0 : b) * s.H + yy) * s.W + xx) * s.c_pitch;
      for (int c = 0; c < s.C; ++c)
        acc = s.fp16 ? fmaf(__half2float(reinterpret_cast<const __half*>(a)[c]),
                            __half2float(reinterpret_cast<const __half*>(wrow)[k + c]), acc)
                     : fmaf(__bfloat162float(a[c]), __bfloat162float(wrow[k + c]), acc);
    }
    k += s.C;
  }
  float v = acc * p.alpha;
  if (p.bias) v += p.bias[n];
  if (p.vec) v += p.vec[(long long)b * p.vec_stride + n];
  if (p.vec2) v += p.vec2[(long long)b * p.vec2_stride + n];
  const long long opix = ((long long)b * p.Hout + (y * p.out_scale + ph.py)) * p.Wout + (x * p.out_scale + ph.px);
  if (p.res)
    v += p.res_fp16 ? __half2float(reinterpret_cast<const __half*>(p.res)[opix * p.res_c_pitch + n])
                    : __bfloat162float(p.res[opix * p.res_c_pitch + n]);
  if (p.out_nchw)
    static_cast<float*>(p.out)[(((long long)b * p.Cout + n) * p.Hout + (y * p.out_scale + ph.py)) * p.Wout +
                               (x * p.out_scale + ph.px)] = v;
  else if (p.out_fp32)
    static_cast<float*>(p.out)[opix * p.out_c_pitch + n] = v;
  else
    if (p.out_fp16)
      static_cast<__half*>(p.out)[opix * p.out_c_pitch + n] = __float2half_rn(v);
    else
      static_cast<__nv_bfloat16*>(p.out)[opix * p.out_c_pitch + n] = __float2bfloat16_rn(v);
}

int tapgemm_launch_ref(const TapGemmParams& p, cudaStream_t stream) {
  const long long rows = (long long)p.B * p.Hm * p.Wm;
  ITS_REQUIRE(rows <= 2147483647LL, "tapgemm_ref: too many rows");
  dim3 grid((unsigned)rows, (p.Cout + 127) / 128, p.nphases);
  ITS_LAUNCH(tapgemm_ref_kernel, dim3(grid), dim3(128), 0, stream, p);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

int tapgemm_build_params(const its_conv_desc* d, TapGemmParams* p, bool need_k64) {
  ITS_REQUIRE(d != nullptr, "its_conv_igemm: null descriptor");
  ITS_REQUIRE(d->nsrc >= 1 && d->nsrc <= ITS_MAX_SRC, "its_conv_igemm: nsrc=%d", d->nsrc);
  ITS_REQUIRE(d->nphases >= 1 && d->nphases <= ITS_MAX_PHASES, "its_conv_igemm: nphases=%d", d->nphases);
  ITS_REQUIRE(d->B > 0 && d->Hm > 0 && d->Wm > 0 && d->Cout > 0, "its_conv_igemm: bad B/Hm/Wm/Cout");
  ITS_REQUIRE(d->w && d->out, "its_conv_igemm: null weight/out pointer");
  ITS_REQUIRE(d->out_scale == 1 || d->out_scale == 2, "its_conv_igemm: out_scale=%d", d->out_scale);
  ITS_REQUIRE(d->Hout == d->Hm * d->out_scale && d->Wout == d->Wm * d->out_scale,
              "its_conv_igemm: Hout/Wout must equal Hm/Wm * out_scale");
  memset(p, 0, sizeof(*p));
  p->nsrc = d->nsrc;
  p->nphases = d->nphases;
  for (int s = 0; s < d->nsrc; ++s) {
    const its_src_t& in = d->src[s];
    ITS_REQUIRE(in.ptr && in.C > 0 && in.c_pitch >= in.c_off + in.C, "its_conv_igemm: src %d channels", s);
    ITS_REQUIRE(in.stride == 1 || in.stride == 2, "its_conv_igemm: src %d stride=%d", s, in.stride);
    ITS_REQUIRE(in.c_pitch % 8 == 0 && in.c_off % 8 == 0, "its_conv_igemm: src %d channel pitch/offset must be multiples of 8", s);
    if (need_k64) ITS_REQUIRE(in.C % 64 == 0, "its_conv_igemm: src %d C=%d not a multiple of 64 (use impl=1)", s, in.C);
    DevSrc& o = p->src[s];
    o.ptr = static_cast<const __nv_bfloat16*>(in.ptr) + in.c_off;
    o.c_pitch = in.c_pitch; o.C = in.C; o.H = in.H; o.W = in.W; o.stride = in.stride; o.bcast = in.bcast; o.fp16 = in.fp16;
  }
  for (int f = 0; f < d->nphases; ++f) {
    const its_phase_t& in = d->phase[f];
    ITS_REQUIRE(in.ntaps >= 1 && in.ntaps <= ITS_MAX_TAPS, "its_conv_igemm: phase %d ntaps=%d", f, in.ntaps);
    DevPhase& o = p->phase[f];
    o.ntaps = in.ntaps; o.w_k0 = in.w_k0; o.py = in.py; o.px = in.px;
    ITS_REQUIRE(in.py >= 0 && in.py < d->out_scale && in.px >= 0 && in.px < d->out_scale, "its_conv_igemm: phase %d offset", f);
    int k = 0;
    for (int t = 0; t < in.ntaps; ++t) {
      ITS_REQUIRE(in.src[t] >= 0 && in.src[t] < d->nsrc, "its_conv_igemm: phase %d tap %d src", f, t);
      o.src[t] = in.src[t]; o.dy[t] = in.dy[t]; o.dx[t] = in.dx[t];
      k += d->src[in.src[t]].C;
    }
    o.nkb = k / 64;
    // operand format along K: taps are grouped by format (GroupNorm-output taps first, raw taps
    // after), so one switch point describes the phase
    o.fp16_first = d->src[in.src[0]].fp16 != 0;
    o.kb_switch = o.nkb;
    {
      int kk = 0, switches = 0;
      for (int t = 0; t < in.ntaps; ++t) {
        const int f = d->src[in.src[t]].fp16 != 0;
        const int prev = (t == 0) ? o.fp16_first : (d->src[in.src[t - 1]].fp16 != 0);
        if (f != prev) {
          ++switches;
          o.kb_switch = kk / 64;
        }
        kk += d->src[in.src[t]].C;
      }
      ITS_REQUIRE(switches <= 1, "its_conv_igemm: phase %d interleaves fp16 and bf16 taps (group them)", f);
    }
    ITS_REQUIRE(in.w_k0 >= 0 && in.w_k0 + k <= d->w_pitch, "its_conv_igemm: phase %d K range [%d,%d) exceeds w_pitch=%d", f, in.w_k0, in.w_k0 + k, d->w_pitch);
    if (need_k64) ITS_REQUIRE(in.w_k0 % 8 == 0, "its_conv_igemm: phase %d w_k0 alignment", f);
  }
  p->B = d->B; p->Hm = d->Hm; p->Wm = d->Wm;
  p->w = static_cast<const __nv_bfloat16*>(d->w);
  p->w_pitch = d->w_pitch; p->w_batch_stride = d->w_batch_stride; p->Cout = d->Cout;
  ITS_REQUIRE(d->out_nchw || d->out_c_pitch >= d->out_c_off + d->Cout, "its_conv_igemm: out channel pitch");
  ITS_REQUIRE(!d->out_nchw || (d->out_fp32 && d->out_c_off == 0 && d->res == nullptr),
              "its_conv_igemm: out_nchw needs fp32 output, no channel offset, no residual");
  p->out_nchw = d->out_nchw;
  p->splits = d->splits > 1 ? d->splits : 1;
  p->ws = d->ws;
  p->dbg = reinterpret_cast<long long*>(d->dbg);
  p->stats = d->stats;
  p->stats_parts = d->stats_parts;
  p->out_fp16 = d->out_fp16 & 1;            // bit 0: output format, bit 1: residual tensor format
  p->res_fp16 = (d->out_fp16 >> 1) & 1;
  ITS_REQUIRE(!((d->out_fp16 & 1) && (d->out_fp32 || d->out_nchw)), "its_conv_igemm: out_fp16 with an fp32 output");
  if (p->splits > 1) ITS_REQUIRE(d->ws != nullptr, "its_conv_igemm: splits=%d needs a workspace", d->splits);
  p->out_fp32 = d->out_fp32;
  p->out = d->out_fp32 ? static_cast<void*>(static_cast<float*>(d->out) + d->out_c_off)
                       : static_cast<void*>(static_cast<__nv_bfloat16*>(d->out) + d->out_c_off);
  p->Hout = d->Hout; p->Wout = d->Wout; p->out_scale = d->out_scale; p->out_c_pitch = d->out_c_pitch;
  p->bias = d->bias;
  p->vec = d->vec ? d->vec + d->vec_off : nullptr;
  p->vec_stride = d->vec_stride;
  p->vec2 = d->vec2 ? d->vec2 + d->vec2_off : nullptr;
  p->vec2_stride = d->vec2_stride;
  p->res = d->res ? static_cast<const __nv_bfloat16*>(d->res) + d->res_c_off : nullptr;
  p->res_c_pitch = d->res_c_pitch;
  p->alpha = d->alpha;
  p->gn_out = d->gn_out; p->gn_c_pitch = d->gn_c_pitch; p->gn_gamma = d->gn_gamma; p->gn_beta = d->gn_beta;
  p->gn_groups = d->gn_groups; p->gn_eps = d->gn_eps; p->gn_silu = d->gn_silu; p->gn_only = d->gn_only;
  p->gn_sync = d->gn_sync;
  // M tiling (used by the tcgen05 kernel): box of 128 GEMM rows
  int bw = d->Wm < 128 ? d->Wm : 128;
  int bh = 128 / bw; if (bh > d->Hm) bh = d->Hm;
  int bb = 128 / (bw * bh);
  if (d->w_batch_stride != 0 && bb > 1) {
    // per-image B operand: a tile must not span images.  Let the box overhang the
    // image instead (TMA zero-fills, the epilogue predicates the rows away).
    bh = 128 / bw;
    bb = 1;
  }
  p->bw = bw; p->bh = bh; p->bb = bb;
  p->tiles_x = (d->Wm + bw - 1) / bw;
  p->tiles_y = (d->Hm + bh - 1) / bh;
  p->tiles_b = (d->B + bb - 1) / bb;
  return ITS_OK;
}

}  // namespace its

extern "C" int its_conv_head(void* out, const float* x, const float* W, const float* bias,
                             int32_t n_img, int32_t n_img_in, int32_t H, int32_t Wd, int32_t Cin,
                             int32_t Cout, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && x && W, "its_conv_head: null pointer");
  ITS_REQUIRE(n_img > 0 && n_img_in > 0 && H > 0 && Wd > 0 && Cin == 3 && Cout % 16 == 0 && Cout > 0,
              "its_conv_head: unsupported shape Cin=%d Cout=%d (Cin must be 3, Cout a multiple of 16)", Cin, Cout);
  const size_t smem = ((size_t)Cin * 9 + 1) * Cout * sizeof(float);
  ITS_REQUIRE(smem <= 48 * 1024, "its_conv_head: Cout=%d too large", Cout);
  const long long npix = (long long)n_img * H * Wd;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ITS_LAUNCH(conv_head_kernel<3>, dim3((unsigned)blocks), dim3(256), smem, as_stream(stream), 
      static_cast<__nv_bfloat16*>(out), x, W, bias, n_img, n_img_in, H, Wd, Cout);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}


extern "C" int its_head_patches(void* out, const float* x, int32_t n_img, int32_t n_img_in, int32_t H,
                                int32_t Wd, int32_t Cin, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && x, "its_head_patches: null pointer");
  ITS_REQUIRE(n_img > 0 && n_img_in > 0 && H > 0 && Wd > 0 && Cin >= 1 && Cin * 27 <= 128,
              "its_head_patches: unsupported shape Cin=%d (3 * 9 * Cin must fit 128 patch channels)", Cin);
  const long long npix = (long long)n_img * H * Wd;
  long long blocks = (npix + HP_THREADS - 1) / HP_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
#define ITS_HP_CASE(CIN_)                                                                              \
  case CIN_:                                                                                         \
    ITS_LAUNCH(head_patches_kernel<CIN_>, dim3((unsigned)blocks), dim3(HP_THREADS), 0, as_stream(stream), \
               static_cast<__nv_bfloat16*>(out), x, npix, n_img_in, H, Wd);                          \
    break;
  switch (Cin) {
    ITS_HP_CASE(1) ITS_HP_CASE(2) ITS_HP_CASE(3) ITS_HP_CASE(4)
    default: return set_error(ITS_ERR_INVALID, "its_head_patches: Cin=%d", Cin);
  }
#undef ITS_HP_CASE
  return ITS_OK;
}

extern "C" int its_conv_tail(float* out, const void* act, const float* W, const float* bias,
                             int32_t n_img, int32_t H, int32_t Wd, int32_t Cin, int32_t Cout,
                             void* stream) {
  using namespace its;
  ITS_REQUIRE(out && act && W, "its_conv_tail: null pointer");
  ITS_REQUIRE(n_img > 0 && H > 0 && Wd > 0 && Cin % 8 == 0 && Cin > 0 && Cout >= 1 && Cout <= 4,
              "its_conv_tail: unsupported shape Cin=%d Cout=%d", Cin, Cout);
  const size_t smem = (size_t)9 * Cin * 4 * sizeof(float);
  ITS_REQUIRE(smem <= 48 * 1024, "its_conv_tail: Cin=%d too large", Cin);
  const long long npix = (long long)n_img * H * Wd;
  long long blocks = (npix * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ITS_LAUNCH(conv_tail_kernel, dim3((unsigned)blocks), dim3(256), smem, as_stream(stream), 
      out, static_cast<const __nv_bfloat16*>(act), W, bias, n_img, H, Wd, Cin, Cout);
  ITS_CHECK_LAUNCH();
  return ITS_OK;
}

extern "C" int its_conv_igemm(const its_conv_desc* desc_host, int32_t impl, void* stream) {
  using namespace its;
  TapGemmParams p;
  int rc = tapgemm_build_params(desc_host, &p, impl == 0);
  if (rc != ITS_OK) return rc;
  ITS_REQUIRE(desc_host->schedule >= 0 && desc_host->schedule <= 2, "its_conv_igemm: schedule=%d", desc_host->schedule);
  ITS_REQUIRE(p.gn_out == nullptr || (impl == 0 && desc_host->schedule != 1),
              "its_conv_igemm: the fused GroupNorm epilogue needs the persistent tcgen05 schedule");
  if (impl == 1) {
    ITS_REQUIRE(p.stats == nullptr, "its_conv_igemm: GroupNorm statistics need the persistent tcgen05 schedule");
    return tapgemm_launch_ref(p, as_stream(stream));
  }
  ITS_REQUIRE(impl == 0, "its_conv_igemm: impl=%d", impl);
  const bool persist = desc_host->schedule != 1 && tapgemm_persist_eligible(desc_host, p);
  ITS_REQUIRE(persist || desc_host->schedule != 2, "its_conv_igemm: schedule=2 but the layer is not eligible for the persistent kernel");
  if (persist) return tapgemm_launch_persist(desc_host, p, as_stream(stream));
  ITS_REQUIRE(p.stats == nullptr, "its_conv_igemm: GroupNorm statistics need the persistent tcgen05 schedule");
  ITS_REQUIRE(p.gn_out == nullptr, "its_conv_igemm: the fused GroupNorm epilogue needs the persistent tcgen05 schedule");
  return tapgemm_launch_sm100(desc_host, p, as_stream(stream));
}

extern "C" int its_conv_gn_sync_words(const its_conv_desc* desc_host) {
  using namespace its;
  TapGemmParams p;
  if (tapgemm_build_params(desc_host, &p, true) != ITS_OK) return -1;
  if (desc_host->schedule == 1 || !tapgemm_persist_eligible(desc_host, p)) return -1;
  return tapgemm_gn_sync_words(desc_host, p);
}

extern "C" int its_conv_stats_parts(const its_conv_desc* desc_host) {
  using namespace its;
  TapGemmParams p;
  if (tapgemm_build_params(desc_host, &p, true) != ITS_OK) return 0;
  if (desc_host->schedule == 1 || !tapgemm_persist_eligible(desc_host, p)) return 0;
  return tapgemm_stats_parts(desc_host, p);
}
