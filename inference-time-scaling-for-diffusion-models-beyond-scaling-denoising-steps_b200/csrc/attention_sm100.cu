// Fused single-head attention core on tcgen05 for feature maps of N = 256 tokens
// (Model.py:153-158 / ModelCondition.py:108-113):
//
//     S = Q K^T            tcgen05.mma, fp32 in TMEM          (128 queries x 256 keys per CTA)
//     P = exp(scale*(S - rowmax))                              (registers, bf16 into swizzled smem)
//     O = (P V) / rowsum + b_v                                 tcgen05.mma over the same TMEM columns
//
// One CTA per (image, 128-query tile).  Q|K come from the fused q,k projection tensor
// [B][N][2C] (bf16, NHWC), V^T from the [B][C][N] tensor the V projection writes with the
// weights as the A operand, so every MMA operand is K-major and loaded by TMA; the score
// matrix and the probabilities never touch global memory.  b_v is added after the product
// because the rows of softmax(S) sum to one.
//
// Warp roles: warp 0 = TMA producer (C/64 stages of {Q panel, K panel}, then N/64 stages of
// V^T panels through one 3-slot ring), warp 1 = MMA issuer, warps 2..9 = softmax + output
// (thread = (query row, column half)).
#include "tapgemm.cuh"
#include "sm100_ptx.cuh"

namespace its {

constexpr int AT_THREADS = 320;
constexpr int AT_EPI = 256;
constexpr int AT_N = 256;                 // tokens (keys) per image
constexpr int AT_STAGE = 48 * 1024;       // Q panel 16 KB + K panel 32 KB, or one V^T panel (<= 48 KB)
constexpr int AT_STAGES = 3;
constexpr int AT_P_OFF = AT_STAGES * AT_STAGE;          // P: 4 panels x 16 KB
constexpr int AT_X_OFF = AT_P_OFF + 4 * 16384;          // row max and row sum exchange: 2 x [2][256] floats
constexpr int AT_BAR_OFF = AT_X_OFF + 4 * 256 * 4;
constexpr int AT_SMEM = AT_BAR_OFF + 16 * 8 + 16;

struct AttnParams {
  const float* bias_v;   // [C] or null
  int C;                 // channels (head dim), multiple of 64, <= 384
  float scale_log2e;     // C^-0.5 * log2(e)
  int v_mn;              // 1: V comes MN-major from the fused q|k|v tensor (tmV boxes of 64 channels x 64 keys)
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_fused_kernel(const AttnParams p, const __grid_constant__ CUtensorMap tmQ,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmO) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* p_smem = smem + AT_P_OFF;
  float* xch = reinterpret_cast<float*>(smem + AT_X_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + AT_BAR_OFF);
  uint64_t* empty_bar = full_bar + AT_STAGES;
  uint64_t* s_full = empty_bar + AT_STAGES;     // S accumulated
  uint64_t* p_ready = s_full + 1;               // P staged, S columns free
  uint64_t* o_full = p_ready + 1;               // O accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, img = blockIdx.y;
  const int C = p.C;
  const int nkc = C / 64;                       // k-blocks of the score GEMM
  constexpr int NKN = AT_N / 64;                // k-blocks of the P V GEMM
  const int nd = (C > 256) ? 2 : 1;             // the output's N extent is split into nd MMAs
  const int dn = C / nd;
  const uint32_t tmem_cols = (C > 256) ? 512u : 256u;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < AT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, AT_EPI);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ----
    uint32_t it = 0;
    for (int kb = 0; kb < nkc + NKN; ++kb, ++it) {
      const uint32_t stage = it % AT_STAGES, parity = (it / AT_STAGES) & 1u;
      mbar_wait(&empty_bar[stage], parity ^ 1u);
      if (elect_one_sync()) {
        uint8_t* dst = smem + stage * AT_STAGE;
        if (kb < nkc) {
          mbar_expect_tx(&full_bar[stage], 16384u + 32768u);
          tma_load_3d(dst, &tmQ, &full_bar[stage], kb * 64, qt * 128, img);
          tma_load_3d(dst + 16384, &tmK, &full_bar[stage], C + kb * 64, 0, img);
        } else {
          const int kn = kb - nkc;               // 64 keys x C channels of V^T
          mbar_expect_tx(&full_bar[stage], (uint32_t)(C * 128));
          if (p.v_mn) {                          // V rows as they are: panels of 64 channels x 64 keys
            for (int cp = 0; cp < nkc; ++cp)
              tma_load_3d(dst + cp * 8192, &tmV, &full_bar[stage], 2 * C + cp * 64, kn * 64, img);
          } else {
            for (int h = 0; h < nd; ++h)
              tma_load_3d(dst + h * dn * 128, &tmV, &full_bar[stage], kn * 64, h * dn, img);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -----
    uint32_t it = 0;
    const uint32_t idesc_s = make_idesc(AT_N);
    const uint32_t idesc_o = make_idesc(dn);
    for (int kb = 0; kb < nkc; ++kb, ++it) {
      const uint32_t stage = it % AT_STAGES, parity = (it / AT_STAGES) & 1u;
      mbar_wait(&full_bar[stage], parity);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_addr = smem_u32(smem + stage * AT_STAGE);
        const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(a_addr + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, (uint32_t)((kb | k) != 0));
        umma_commit(&empty_bar[stage]);
        if (kb == nkc - 1) umma_commit(s_full);
      }
      __syncwarp();
    }
    mbar_wait(p_ready, 0);                       // P is in shared memory, S has been read out of TMEM
    tcgen05_fence_after();
    for (int kn = 0; kn < NKN; ++kn, ++it) {
      const uint32_t stage = it % AT_STAGES, parity = (it / AT_STAGES) & 1u;
      mbar_wait(&full_bar[stage], parity);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint64_t adesc = make_smem_desc(smem_u32(p_smem + kn * 16384));
        const uint32_t b_addr = smem_u32(smem + stage * AT_STAGE);
        for (int h = 0; h < nd; ++h) {
          if (p.v_mn) {
            // B = V[64 keys][dn channels], MN-major: channel panels 8 KB apart, 16 keys = 2 KB further on
            const uint64_t bdesc = make_smem_desc_mn(b_addr + h * (dn / 64) * 8192, 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + h * dn, adesc + 2 * k, bdesc + 128 * k, idesc_o | IDESC_B_MN_MAJOR,
                        (uint32_t)((kn | k) != 0));
          } else {
            const uint64_t bdesc = make_smem_desc(b_addr + h * dn * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + h * dn, adesc + 2 * k, bdesc + 2 * k, idesc_o, (uint32_t)((kn | k) != 0));
          }
        }
        umma_commit(&empty_bar[stage]);
        if (kn == NKN - 1) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------- softmax + output -----
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(s_full, 0);
    tcgen05_fence_after();
    // pass 1: row maximum over this thread's 128 keys
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32_nowait(lane_addr + (uint32_t)(half * 128 + c * 32), v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    xch[half * 256 + row] = mx;
    named_bar_sync(1, AT_EPI);
    mx = fmaxf(mx, xch[(half ^ 1) * 256 + row]);
    const float moff = mx * p.scale_log2e;
    // pass 2: p = exp2(scale*log2e*s - scale*log2e*max), bf16 into the swizzled A-operand panels
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32_nowait(lane_addr + (uint32_t)(half * 128 + c * 32), v);
      tmem_wait_ld();
      const int key0 = half * 128 + c * 32;      // first key of this chunk
      const int panel = key0 >> 6, chunk0 = (key0 & 63) >> 3;
      const uint32_t row_addr = smem_u32(p_smem + panel * 16384) + (uint32_t)row * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = fmaf(__uint_as_float(v[g * 8 + i]), p.scale_log2e, -moff);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(a));
        }
        const bf16x8 pk = pack8(f);
        // the row sum is taken over the bf16-rounded probabilities the MMA will read
        float r[8];
        unpack8(pk, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += r[i];
        const uint32_t dst = row_addr + (uint32_t)(((chunk0 + g) ^ (row & 7)) << 4);
        const uint4 u = *reinterpret_cast<const uint4*>(&pk);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
    }
    xch[512 + half * 256 + row] = sum;           // (second exchange array follows the first)
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(p_ready);
    mbar_wait(o_full, 0);
    tcgen05_fence_after();
    named_bar_sync(1, AT_EPI);                   // partner's partial sum is visible; ring slots are free
    const float inv = 1.0f / (sum + xch[512 + (half ^ 1) * 256 + row]);
    // output: this thread's half of the channels, 32 at a time, staged as 64-channel panels
    const int cols_half = C / 2;                 // multiple of 32
    uint8_t* stg = smem;                         // the operand ring is idle now
#pragma unroll 1
    for (int c = 0; c < cols_half; c += 32) {
      const int col0 = half * cols_half + c;
      uint32_t v[32];
      tmem_ld32_nowait(lane_addr + (uint32_t)col0, v);
      tmem_wait_ld();
      const int panel = col0 >> 6, chunk0 = (col0 & 63) >> 3;
      const uint32_t row_addr = smem_u32(stg + panel * 16384) + (uint32_t)row * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]) * inv;
        if (p.bias_v) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias_v + col0 + g * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias_v + col0 + g * 8 + 4));
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
          f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        const bf16x8 pk = pack8(f);
        const uint32_t dst = row_addr + (uint32_t)(((chunk0 + g) ^ (row & 7)) << 4);
        const uint4 u = *reinterpret_cast<const uint4*>(&pk);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, AT_EPI);
    if (warp == 2 && lane == 0) {
      for (int pn = 0; pn < nkc; ++pn) tma_store_3d(&tmO, stg + pn * 16384, pn * 64, qt * 128, img);
      bulk_commit_group();
      bulk_wait_group<0>();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace its

extern "C" int its_attention_fused(void* out, const void* qk, const void* vT, const float* bias_v,
                                   int32_t n_img, int32_t N, int32_t C, float scale, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && qk, "its_attention_fused: null pointer");
  const bool v_mn = (vT == nullptr);     // qk is the fused q|k|v tensor [n_img][N][3C]
  const int pitch = v_mn ? 3 * C : 2 * C;
  ITS_REQUIRE(N == AT_N, "its_attention_fused: N=%d tokens (only %d supported; use the GEMM + softmax path)", N, AT_N);
  ITS_REQUIRE(C % 64 == 0 && C >= 64 && C <= 384, "its_attention_fused: C=%d must be a multiple of 64 in [64, 384]", C);
  ITS_REQUIRE(n_img > 0, "its_attention_fused: n_img");
  ITS_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(vT)) & 15) == 0,
              "its_attention_fused: pointer alignment");
  ITS_REQUIRE(!v_mn || C <= 256 || (C / 2) % 64 == 0, "its_attention_fused: C=%d", C);
  static_assert(AT_SMEM <= 227 * 1024, "shared memory budget");
  CUtensorMap tmQ, tmK, tmV, tmO;
  const cuuint32_t estr[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)N, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)N * pitch * 2};
    const cuuint32_t boxq[3] = {64, 128, 1};
    const cuuint32_t boxk[3] = {64, (cuuint32_t)AT_N, 1};
    int rc = encode_bf16_map(&tmQ, 3, qk, dims, strides, boxq, estr, "attention Q");
    if (rc != ITS_OK) return rc;
    rc = encode_bf16_map(&tmK, 3, qk, dims, strides, boxk, estr, "attention K");
    if (rc != ITS_OK) return rc;
  }
  if (v_mn) {
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)N, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)N * pitch * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    int rc = encode_bf16_map(&tmV, 3, qk, dims, strides, box, estr, "attention V");
    if (rc != ITS_OK) return rc;
  } else {
    const int dn = C > 256 ? C / 2 : C;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)C * N * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)dn, 1};
    int rc = encode_bf16_map(&tmV, 3, vT, dims, strides, box, estr, "attention V^T");
    if (rc != ITS_OK) return rc;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)N * C * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    int rc = encode_bf16_map(&tmO, 3, out, dims, strides, box, estr, "attention out");
    if (rc != ITS_OK) return rc;
  }
  AttnParams p;
  p.bias_v = bias_v;
  p.C = C;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.v_mn = v_mn ? 1 : 0;
  static PerDeviceBytes configured;
  if (configured.need(AT_SMEM))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(attention_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  ITS_LAUNCH(attention_fused_kernel, dim3(N / 128, n_img), dim3(AT_THREADS), (size_t)AT_SMEM, as_stream(stream), p, tmQ,
             tmK, tmV, tmO);
  return ITS_OK;
}
