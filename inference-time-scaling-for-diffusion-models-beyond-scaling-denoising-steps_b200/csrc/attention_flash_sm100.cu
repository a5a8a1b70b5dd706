// Streaming ("flash") single-head attention core on tcgen05 for feature maps with more
// tokens than one TMEM-resident score tile can hold (ModelCondition.py:108-113 on the 32x32
// maps of the conditional net: N = 1024 tokens, head dimension = channels = 128):
//
//   for each block j of 128 keys:   S_j = Q K_j^T                      tcgen05.mma -> TMEM
//                                   m'  = max(m, rowmax S_j)           softmax warps, registers
//                                   P_j = exp(scale (S_j - m'))        bf16 -> swizzled smem
//                                   l   = l exp(scale (m - m')) + rowsum P_j
//                                   T_j = P_j V_j                      tcgen05.mma -> TMEM (fresh)
//                                   O   = O exp(scale (m - m')) + T_j  output warps, registers
//   out = O / l + b_v
//
// One CTA per (image, 128-query tile).  The score matrix and the probabilities never touch
// global memory (the unfused path writes 4 N^2 bytes of fp32 scores per image and reads them
// back twice).  Unlike the N = 256 kernel (attention_sm100.cu) the running output is kept in
// registers — each block's P_j V_j lands in its own TMEM buffer and is folded in one block later —
// so no TMEM read-modify-write is needed when the running maximum moves.
//
// Warp roles: warp 0 = TMA producer (Q once; then K_0, K_1, V_0, K_2, V_1, ... through a 3-slot
// ring), warp 1 = MMA issuer (S_{j+1} is issued before P_j V_j, so the tensor pipe computes the
// next scores while the softmax warps work on the current ones), warps 2..9 = softmax + output
// accumulation (thread = (query row, half): 64 keys of every block, C/2 fp32 output registers).
// TMEM: S double buffered at columns [0,128) / [128,256); T at [256,256+C) / [384,384+C).
#include "tapgemm.cuh"
#include "sm100_ptx.cuh"

namespace its {

constexpr int FL_THREADS = 320;
constexpr int FL_KB = 128;                 // keys per block
constexpr int FL_STAGES = 3;
constexpr int FL_NBARS = 1 + 2 * FL_STAGES + 12;

template <int C>
struct FlashSmem {
  static constexpr int Q_BYTES = 128 * C * 2;              // C/64 panels of 128 rows x 128 B
  static constexpr int STAGE_BYTES = C * 256;              // K block (C/64 panels x 16 KB) or V^T block (2 x C rows x 128 B)
  static constexpr int P_BYTES = 128 * FL_KB * 2;          // 2 panels x 16 KB
  static constexpr int RING_OFF = Q_BYTES;
  static constexpr int P_OFF = RING_OFF + FL_STAGES * STAGE_BYTES;
  static constexpr int X_OFF = P_OFF + 2 * P_BYTES;        // row-max exchange [2][2][128], row sums [2][128] floats
  static constexpr int BAR_OFF = X_OFF + 6 * 128 * 4;
  static constexpr int TOTAL = BAR_OFF + FL_NBARS * 8 + 16;
};

struct FlashParams {
  const float* bias_v;
  int N;
  float scale_log2e;
  int v_mn;      // 1: V blocks come MN-major from the fused q|k|v tensor (boxes of 64 channels x 128 keys)
};

template <int C>
__global__ void __launch_bounds__(FL_THREADS, 1)
attention_flash_kernel(const FlashParams p, const __grid_constant__ CUtensorMap tmQ,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __grid_constant__ CUtensorMap tmO) {
  using L = FlashSmem<C>;
  static_assert(C == 64 || C == 128, "head dimension");
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + L::RING_OFF;
  uint8_t* p_smem = smem + L::P_OFF;
  float* alpha_s = reinterpret_cast<float*>(smem + L::X_OFF);    // row-max exchange [2 buffers][2 halves][128]
  float* l_s = alpha_s + 4 * 128;                                // partial row sums [2 halves][128]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* full_bar = q_full + 1;
  uint64_t* empty_bar = full_bar + FL_STAGES;
  uint64_t* s_full = empty_bar + FL_STAGES;    // [2] scores of a block accumulated
  uint64_t* s_free = s_full + 2;               // [2] scores read out of TMEM
  uint64_t* p_full = s_free + 2;               // [2] probabilities (and alpha) staged
  uint64_t* p_free = p_full + 2;               // [2] P V has consumed the probability buffer
  uint64_t* t_full = p_free + 2;               // [2] P V accumulated
  uint64_t* t_free = t_full + 2;               // [2] P V read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, img = blockIdx.y;
  const int NB = p.N / FL_KB;
  constexpr int NKC = C / 64;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FL_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], 256);
      mbar_init(&p_full[b], 256);
      mbar_init(&p_free[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_free[b], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
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
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, (uint32_t)L::Q_BYTES);
      for (int kc = 0; kc < NKC; ++kc) tma_load_3d(q_smem + kc * 16384, &tmQ, q_full, kc * 64, qt * 128, img);
    }
    __syncwarp();
    uint32_t it = 0;
    for (int j = 0; j <= NB; ++j) {
      if (j < NB) {                                // K_j: 128 keys x C channels
        const uint32_t stage = it % FL_STAGES, parity = (it / FL_STAGES) & 1u;
        mbar_wait(&empty_bar[stage], parity ^ 1u);
        if (elect_one_sync()) {
          uint8_t* dst = ring + stage * L::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)L::STAGE_BYTES);
          for (int kc = 0; kc < NKC; ++kc)
            tma_load_3d(dst + kc * 16384, &tmK, &full_bar[stage], C + kc * 64, j * FL_KB, img);
        }
        __syncwarp();
        ++it;
      }
      if (j >= 1) {                                // V^T_{j-1}: C channels x 128 keys (two 64-key panels)
        const uint32_t stage = it % FL_STAGES, parity = (it / FL_STAGES) & 1u;
        mbar_wait(&empty_bar[stage], parity ^ 1u);
        if (elect_one_sync()) {
          uint8_t* dst = ring + stage * L::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)L::STAGE_BYTES);
          if (p.v_mn) {
            for (int cp = 0; cp < NKC; ++cp)
              tma_load_3d(dst + cp * 16384, &tmV, &full_bar[stage], 2 * C + cp * 64, (j - 1) * FL_KB, img);
          } else {
            for (int kp = 0; kp < 2; ++kp)
              tma_load_3d(dst + kp * C * 128, &tmV, &full_bar[stage], (j - 1) * FL_KB + kp * 64, 0, img);
          }
        }
        __syncwarp();
        ++it;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -----
    constexpr uint32_t idesc_s = make_idesc(FL_KB);
    constexpr uint32_t idesc_t = make_idesc(C);
    mbar_wait(q_full, 0);
    uint32_t it = 0;
    for (int j = 0; j <= NB; ++j) {
      if (j < NB) {
        const uint32_t b = j & 1u, use = (uint32_t)(j >> 1);
        const uint32_t stage = it % FL_STAGES, parity = (it / FL_STAGES) & 1u;
        mbar_wait(&s_free[b], (use & 1u) ^ 1u);      // softmax has read S_{j-2}
        mbar_wait(&full_bar[stage], parity);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t k_addr = smem_u32(ring + stage * L::STAGE_BYTES);
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc) {
            const uint64_t adesc = make_smem_desc(smem_u32(q_smem + kc * 16384));
            const uint64_t bdesc = make_smem_desc(k_addr + kc * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + b * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, (uint32_t)((kc | k) != 0));
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&s_full[b]);
        }
        __syncwarp();
        ++it;
      }
      if (j >= 1) {
        const int i = j - 1;
        const uint32_t b = i & 1u, use = (uint32_t)(i >> 1);
        const uint32_t stage = it % FL_STAGES, parity = (it / FL_STAGES) & 1u;
        mbar_wait(&p_full[b], use & 1u);             // P_i staged
        mbar_wait(&t_free[b], (use & 1u) ^ 1u);      // T_{i-2} read out
        mbar_wait(&full_bar[stage], parity);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t v_addr = smem_u32(ring + stage * L::STAGE_BYTES);
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {
            const uint64_t adesc = make_smem_desc(smem_u32(p_smem + b * L::P_BYTES + kp * 16384));
            if (p.v_mn) {
              // B = V[128 keys][C channels], MN-major: channel panels 16 KB apart, 16 keys = 2 KB further on
              const uint64_t bdesc = make_smem_desc_mn(v_addr, 16384);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + 256 + b * 128, adesc + 2 * k, bdesc + 128 * (kp * 4 + k),
                          idesc_t | IDESC_B_MN_MAJOR, (uint32_t)((kp | k) != 0));
            } else {
              const uint64_t bdesc = make_smem_desc(v_addr + kp * C * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + 256 + b * 128, adesc + 2 * k, bdesc + 2 * k, idesc_t, (uint32_t)((kp | k) != 0));
            }
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&p_free[b]);
          umma_commit(&t_full[b]);
        }
        __syncwarp();
        ++it;
      }
    }
  } else {
    // ------------------------------- softmax + output accumulation -----
    // thread = (query row, half): the 64 keys [half*64, half*64+64) of every block for the softmax and
    // the C/2 channels [half*C/2, ...) of the running output.  Two threads per row keep two warps per
    // scheduler busy (one warp per scheduler left the MUFU / issue latency exposed: 4000 clocks per
    // block against 1570 of tensor work); the row maximum is exchanged through shared memory.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int CH = C / 2;                     // output channels per thread
    float o[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    for (int j = 0; j < NB; ++j) {
      const uint32_t b = j & 1u, use = (uint32_t)(j >> 1);
      mbar_wait(&s_full[b], use & 1u);
      tcgen05_fence_after();
      // this thread's 64 scores of the block stay in registers between the maximum and the exponentials
      uint32_t v[64];
      tmem_ld32_nowait(lane_addr + b * 128 + (uint32_t)(half * 64), v);
      tmem_ld32_nowait(lane_addr + b * 128 + (uint32_t)(half * 64 + 32), v + 32);
      tmem_wait_ld();
      float mx = m;
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      alpha_s[(b * 2 + half) * 128 + row] = mx;
      named_bar_sync(2, 256);
      mx = fmaxf(mx, alpha_s[(b * 2 + (half ^ 1)) * 128 + row]);
      const float moff = mx * p.scale_log2e;
      float alpha;
      {
        const float a = fmaf(m, p.scale_log2e, -moff);       // -inf at j = 0 -> alpha = 0
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(a));
      }
      mbar_wait(&p_free[b], (use & 1u) ^ 1u);                // P V_{j-2} has consumed this buffer
      float sum = 0.f;
      {
        const uint32_t row_addr = smem_u32(p_smem + b * L::P_BYTES + half * 16384) + (uint32_t)row * 128u;   // panel = half
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a = fmaf(__uint_as_float(v[g * 8 + i]), p.scale_log2e, -moff);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f[i]) : "f"(a));
          }
          const bf16x8 pk = pack8(f);
          float r[8];
          unpack8(pk, r);                  // the row sum is taken over the rounded values the MMA reads
#pragma unroll
          for (int i = 0; i < 8; ++i) sum += r[i];
          const uint32_t dst = row_addr + (uint32_t)((g ^ (row & 7)) << 4);
          const uint4 u = *reinterpret_cast<const uint4*>(&pk);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                       : "memory");
        }
      }
      l = fmaf(l, alpha, sum);
      m = mx;
      tcgen05_fence_before();
      mbar_arrive(&s_free[b]);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[b]);
      if (j >= 1) {
        // deferred by one block: T_{j-1} = P_{j-1} V_{j-1} has had the whole softmax of block j to finish
        const uint32_t bt = (j - 1) & 1u, uset = (uint32_t)((j - 1) >> 1);
        mbar_wait(&t_full[bt], uset & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < CH / 32; ++c) {
          uint32_t v[32];
          tmem_ld32_nowait(lane_addr + 256u + bt * 128 + (uint32_t)(half * CH + c * 32), v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
        tcgen05_fence_before();
        mbar_arrive(&t_free[bt]);
      }
      alpha_prev = alpha;
    }
    {
      const uint32_t bt = (NB - 1) & 1u, uset = (uint32_t)((NB - 1) >> 1);
      mbar_wait(&t_full[bt], uset & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < CH / 32; ++c) {
        uint32_t v[32];
        tmem_ld32_nowait(lane_addr + 256u + bt * 128 + (uint32_t)(half * CH + c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
      }
    }
    l_s[half * 128 + row] = l;
    named_bar_sync(2, 256);
    const float inv = 1.0f / (l + l_s[(half ^ 1) * 128 + row]);
    // every MMA has completed: the Q panels are free and become the output staging
    uint8_t* stg = q_smem;
#pragma unroll
    for (int g = 0; g < CH / 8; ++g) {
      const int col0 = half * CH + g * 8;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = o[g * 8 + i] * inv;
      if (p.bias_v) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias_v + col0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias_v + col0 + 4));
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      const bf16x8 pk = pack8(f);
      const int panel = col0 >> 6, chunk = (col0 & 63) >> 3;
      const uint32_t dst = smem_u32(stg + panel * 16384) + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
      const uint4 u = *reinterpret_cast<const uint4*>(&pk);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                   : "memory");
    }
    fence_proxy_async_smem();
    named_bar_sync(2, 256);
    if (warp == 2 && lane == 0) {
      for (int pn = 0; pn < NKC; ++pn) tma_store_3d(&tmO, stg + pn * 16384, pn * 64, qt * 128, img);
      bulk_commit_group();
      bulk_wait_group<0>();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int C>
static int launch_flash(const FlashParams& p, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                        const CUtensorMap& tmO, int n_img, cudaStream_t stream) {
  using L = FlashSmem<C>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static PerDeviceBytes configured;
  if (configured.need(L::TOTAL))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(attention_flash_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  ITS_LAUNCH(attention_flash_kernel<C>, dim3(p.N / 128, n_img), dim3(FL_THREADS), (size_t)L::TOTAL, stream, p, tmQ, tmK,
             tmV, tmO);
  return ITS_OK;
}

}  // namespace its

extern "C" int its_attention_flash(void* out, const void* qk, const void* vT, const float* bias_v,
                                   int32_t n_img, int32_t N, int32_t C, float scale, void* stream) {
  using namespace its;
  ITS_REQUIRE(out && qk, "its_attention_flash: null pointer");
  const bool v_mn = (vT == nullptr);     // qk is the fused q|k|v tensor [n_img][N][3C]
  const int pitch = v_mn ? 3 * C : 2 * C;
  ITS_REQUIRE(N % FL_KB == 0 && N >= 2 * FL_KB, "its_attention_flash: N=%d tokens must be a multiple of %d, at least %d", N,
              FL_KB, 2 * FL_KB);
  ITS_REQUIRE(C == 64 || C == 128, "its_attention_flash: C=%d (64 or 128 supported; use the GEMM + softmax path)", C);
  ITS_REQUIRE(n_img > 0 && n_img <= 65535, "its_attention_flash: n_img");
  ITS_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(vT)) & 15) == 0,
              "its_attention_flash: pointer alignment");
  CUtensorMap tmQ, tmK, tmV, tmO;
  const cuuint32_t estr[3] = {1, 1, 1};
  {
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)N, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)N * pitch * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    int rc = encode_bf16_map(&tmQ, 3, qk, dims, strides, box, estr, "attention Q");
    if (rc != ITS_OK) return rc;
    rc = encode_bf16_map(&tmK, 3, qk, dims, strides, box, estr, "attention K");
    if (rc != ITS_OK) return rc;
  }
  if (v_mn) {
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)N, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)N * pitch * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    int rc = encode_bf16_map(&tmV, 3, qk, dims, strides, box, estr, "attention V");
    if (rc != ITS_OK) return rc;
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)n_img};
    const cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)C * N * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)C, 1};
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
  FlashParams p;
  p.bias_v = bias_v;
  p.N = N;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.v_mn = v_mn ? 1 : 0;
  if (C == 64) return launch_flash<64>(p, tmQ, tmK, tmV, tmO, n_img, as_stream(stream));
  return launch_flash<128>(p, tmQ, tmK, tmV, tmO, n_img, as_stream(stream));
}
