// Fused single-head attention core on tcgen05 for SMALL feature maps (N = 16, 32 or 64 tokens: the
// 4x4 / 8x8 maps of Model.py:153-158 and ModelCondition.py:108-113), where one 128-row MMA tile spans
// G = 128 / N whole images:
//
//     S = Q K^T            128 tokens x 128 tokens of G consecutive images        (tcgen05, TMEM)
//     P = softmax over the N keys of the row's OWN image, zero elsewhere           (block-diagonal mask)
//     O = (P V) / rowsum + b_v                                                     (tcgen05, same TMEM columns)
//
// Operands come from the fused q|k|v projection tensor [n_img * N][3C] (bf16, NHWC rows): Q and K
// K-major in 64-channel panels, V as an MN-major B operand (rows = keys), all by TMA.  The cross-image
// blocks of S are computed and discarded: at these sizes the tensor work is negligible, what matters
// is that the whole block is one launch at tensor-core speed instead of a CUDA-core loop.
// Rows past the last image (ragged final group) are TMA zero fill on load and clipped on store.
//
// Warp roles: warp 0 = TMA producer ({Q panel, K panel} per 64 channels, then V in blocks of 32 keys
// through one ring of 32 KB slots), warp 1 = MMA issuer, warps 2..9 = softmax + output (thread =
// (row, 64-key half)).
#include "tapgemm.cuh"
#include "sm100_ptx.cuh"

namespace its {

constexpr int AG_THREADS = 320;
constexpr int AG_EPI = 256;
constexpr int AG_STAGE = 32 * 1024;       // Q panel 16 KB + K panel 16 KB, or 32 keys x C channels of V (<= 32 KB)
constexpr int AG_STAGES = 4;              // the ring doubles as the output staging: C / 64 panels of 16 KB
constexpr int AG_P_OFF = AG_STAGES * AG_STAGE;          // P: 2 panels x 16 KB
constexpr int AG_X_OFF = AG_P_OFF + 2 * 16384;          // row max / row sum exchange: 2 x [2][128] floats
constexpr int AG_BAR_OFF = AG_X_OFF + 4 * 128 * 4;
constexpr int AG_SMEM = AG_BAR_OFF + 16 * 8 + 16;

struct AttnGroupParams {
  const float* bias_v;   // [C] or null
  int C;                 // channels (head dim), multiple of 64, <= 512
  int n_shift;           // log2(tokens per image)
  float scale_log2e;
};

__global__ void __launch_bounds__(AG_THREADS, 1)
attention_group_kernel(const AttnGroupParams p, const __grid_constant__ CUtensorMap tmQK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* p_smem = smem + AG_P_OFF;
  float* xch = reinterpret_cast<float*>(smem + AG_X_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + AG_BAR_OFF);
  uint64_t* empty_bar = full_bar + AG_STAGES;
  uint64_t* s_full = empty_bar + AG_STAGES;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;            // first token row of this group of images
  const int C = p.C;
  const int nkc = C / 64;
  const int nd = (C > 256) ? 2 : 1;             // the output's N extent is split into nd MMAs
  const int dn = C / nd;
  const uint32_t tmem_cols = (C > 256) ? 512u : (C > 128) ? 256u : 128u;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < AG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, AG_EPI);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQK);
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
    for (int kb = 0; kb < nkc + 4; ++kb, ++it) {
      const uint32_t stage = it % AG_STAGES, parity = (it / AG_STAGES) & 1u;
      mbar_wait(&empty_bar[stage], parity ^ 1u);
      if (elect_one_sync()) {
        uint8_t* dst = smem + stage * AG_STAGE;
        if (kb < nkc) {
          mbar_expect_tx(&full_bar[stage], 32768u);
          tma_load_3d(dst, &tmQK, &full_bar[stage], kb * 64, row0, 0);
          tma_load_3d(dst + 16384, &tmQK, &full_bar[stage], C + kb * 64, row0, 0);
        } else {
          const int kv = kb - nkc;               // 32 keys x C channels of V, 64-channel panels of 4 KB
          mbar_expect_tx(&full_bar[stage], (uint32_t)(C * 64));
          for (int cp = 0; cp < nkc; ++cp)
            tma_load_3d(dst + cp * 4096, &tmV, &full_bar[stage], 2 * C + cp * 64, row0 + kv * 32, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -----
    uint32_t it = 0;
    const uint32_t idesc_s = make_idesc(128);
    const uint32_t idesc_o = make_idesc(dn) | IDESC_B_MN_MAJOR;
    for (int kb = 0; kb < nkc; ++kb, ++it) {
      const uint32_t stage = it % AG_STAGES, parity = (it / AG_STAGES) & 1u;
      mbar_wait(&full_bar[stage], parity);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t a_addr = smem_u32(smem + stage * AG_STAGE);
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
    for (int kv = 0; kv < 4; ++kv, ++it) {
      const uint32_t stage = it % AG_STAGES, parity = (it / AG_STAGES) & 1u;
      mbar_wait(&full_bar[stage], parity);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        // keys kv*32 .. kv*32+31: P panel kv / 2 (64 keys each), 32-key half kv % 2
        const uint64_t adesc = make_smem_desc(smem_u32(p_smem + (kv >> 1) * 16384)) + (uint64_t)((kv & 1) * 4);
        const uint32_t b_addr = smem_u32(smem + stage * AG_STAGE);
        for (int h = 0; h < nd; ++h) {
          const uint64_t bdesc = make_smem_desc_mn(b_addr + h * (dn / 64) * 4096, 4096);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(tmem_base + h * dn, adesc + 2 * k, bdesc + 128 * k, idesc_o, (uint32_t)((kv | k) != 0));
        }
        umma_commit(&empty_bar[stage]);
        if (kv == 3) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------- softmax + output -----
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int key_lo = (row >> p.n_shift) << p.n_shift, key_hi = key_lo + (1 << p.n_shift);   // own image's keys
    mbar_wait(s_full, 0);
    tcgen05_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      const int key0 = half * 64 + c * 32;
      tmem_ld32_nowait(lane_addr + (uint32_t)key0, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (key0 + i >= key_lo && key0 + i < key_hi) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    xch[half * 128 + row] = mx;
    named_bar_sync(1, AG_EPI);
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + row]);
    const float moff = mx * p.scale_log2e;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      const int key0 = half * 64 + c * 32;
      tmem_ld32_nowait(lane_addr + (uint32_t)key0, v);
      tmem_wait_ld();
      const int chunk0 = c * 4;
      const uint32_t row_addr = smem_u32(p_smem + half * 16384) + (uint32_t)row * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = key0 + g * 8 + i;
          const float a = fmaf(__uint_as_float(v[g * 8 + i]), p.scale_log2e, -moff);
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
          f[i] = (key >= key_lo && key < key_hi) ? e : 0.f;
        }
        const bf16x8 pk = pack8(f);
        float r[8];
        unpack8(pk, r);                  // the row sum is taken over the rounded values the MMA reads
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += r[i];
        const uint32_t dst = row_addr + (uint32_t)(((chunk0 + g) ^ (row & 7)) << 4);
        const uint4 u = *reinterpret_cast<const uint4*>(&pk);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
    }
    xch[256 + half * 128 + row] = sum;
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(p_ready);
    mbar_wait(o_full, 0);
    tcgen05_fence_after();
    named_bar_sync(1, AG_EPI);                   // partner's partial sum is visible; ring slots are free
    const float inv = 1.0f / (sum + xch[256 + (half ^ 1) * 128 + row]);
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
    named_bar_sync(1, AG_EPI);
    if (warp == 2 && lane == 0) {
      for (int pn = 0; pn < nkc; ++pn) tma_store_3d(&tmO, stg + pn * 16384, pn * 64, row0, 0);
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

extern "C" int its_attention_group(void* out, const void* qkv, const float* bias_v, int32_t n_img, int32_t N,
                                   int32_t C, float scale, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && qkv, "its_attention_group: null pointer");
  ITS_REQUIRE(N == 16 || N == 32 || N == 64, "its_attention_group: N=%d tokens (16, 32 or 64)", N);
  ITS_REQUIRE(C % 64 == 0 && C >= 64 && C <= 512, "its_attention_group: C=%d must be a multiple of 64 in [64, 512]", C);
  ITS_REQUIRE(C <= 256 || (C / 2) % 64 == 0, "its_attention_group: C=%d: each output half must be whole 64-channel panels", C);
  ITS_REQUIRE(n_img > 0, "its_attention_group: n_img");
  ITS_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(qkv)) & 15) == 0,
              "its_attention_group: pointer alignment");
  static_assert(AG_SMEM <= 227 * 1024, "shared memory budget");
  static_assert(AG_STAGES * AG_STAGE >= 512 * 256, "the ring must hold the staged output of C = 512");
  const long long rows = (long long)n_img * N;
  CUtensorMap tmQK, tmV, tmO;
  const cuuint32_t estr[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)3 * C, (cuuint64_t)rows, 1};
    const cuuint64_t strides[2] = {(cuuint64_t)3 * C * 2, (cuuint64_t)rows * 3 * C * 2};
    const cuuint32_t boxqk[3] = {64, 128, 1};
    const cuuint32_t boxv[3] = {64, 32, 1};
    int rc = encode_bf16_map(&tmQK, 3, qkv, dims, strides, boxqk, estr, "attention Q/K");
    if (rc != ITS_OK) return rc;
    rc = encode_bf16_map(&tmV, 3, qkv, dims, strides, boxv, estr, "attention V");
    if (rc != ITS_OK) return rc;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, 1};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)rows * C * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    int rc = encode_bf16_map(&tmO, 3, out, dims, strides, box, estr, "attention out");
    if (rc != ITS_OK) return rc;
  }
  AttnGroupParams p;
  p.bias_v = bias_v;
  p.C = C;
  p.n_shift = (N == 16) ? 4 : (N == 32) ? 5 : 6;
  p.scale_log2e = scale * 1.4426950408889634f;
  static PerDeviceBytes configured;
  if (configured.need(AG_SMEM))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(attention_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AG_SMEM));
  const unsigned groups = (unsigned)((rows + 127) / 128);
  ITS_LAUNCH(attention_group_kernel, dim3(groups), dim3(AG_THREADS), (size_t)AG_SMEM, as_stream(stream), p, tmQK, tmV, tmO);
  return ITS_OK;
}
