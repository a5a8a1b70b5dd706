// Persistent tap-GEMM (implicit-GEMM convolution) for sm_100a.
//
// Same maths and operand layout as conv_igemm_sm100.cu (see its_conv_igemm in
// include/its_b200.h), different schedule:
//
// * one CTA per SM walks a static list of output tiles (phase, N tile, M tile), so the
//   barrier/TMEM set-up and the first TMA round trip are paid once per CTA, not per tile;
// * the fp32 accumulator is double buffered in TMEM (2 x BN columns): while the four
//   epilogue warps drain tile j, the MMA warp already accumulates tile j+1 into the other
//   buffer, fed by the same shared-memory ring;
// * the epilogue works on 64-column panels: tcgen05.ld -> alpha/bias/per-image vectors ->
//   bf16 -> 128B-swizzled shared memory -> one TMA tensor store per panel (two panel
//   buffers, bulk-group completion), so no thread issues an uncoalesced global store;
// * GroupNorm statistics of the tensor being written are a by-product: column sums and
//   sums of squares of the bf16 values in the staged panel, per (image, 4 channels, tile),
//   written to a small partial-sum array that the GroupNorm-apply kernel of the consuming
//   layer reduces in a fixed order (Model.py:170,186: GroupNorm(32, C) reads exactly the
//   tensor this kernel stores).
//
// Restrictions (the launcher falls back to the one-tile-per-CTA kernel otherwise):
// bf16 NHWC output, weights shared by all images, no split-K, no residual pointer (the
// identity shortcut is appended as an extra K block with identity weights by the caller),
// Cout a multiple of 64.
#include "tapgemm.cuh"
#include "sm100_ptx.cuh"

namespace its {

constexpr int P_THREADS = 352;       // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 TMA stores
constexpr int P_EPI_THREADS = 256;
constexpr int PANEL_COLS = 64;
constexpr int PANEL_BYTES = BM * PANEL_COLS * 2;   // 128 rows x 128 bytes

// MT = 128-row sub-tiles (accumulators) per tile: MT = 2 shares every weight tile between two
// vertically adjacent row boxes, halving the weight traffic per MMA.
// KS = 64-deep k-blocks per ring stage: with KS = 2 one full/empty barrier round trip (and one pass
// of the MMA warp's wait / elect / commit path, ~200 ns of single-warp issue latency) covers twice
// the MMA work, which is what bounds the narrow N tiles.
template <int BN, int STAGES, int MT, int KS, bool TM = false>
struct PersistSmem {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = KS * (MT * A_BYTES + B_BYTES);
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;          // 2 panel buffers
  static constexpr int RED_OFF = STAGING_OFF + 2 * PANEL_BYTES;     // [8 sub-blocks][16 chunks] float2
  static constexpr int RED_BYTES = 2 * 8 * 16 * 8;   // double buffered
  static constexpr int BAR_OFF = RED_OFF + RED_BYTES;
  static constexpr int NBARS = 2 * STAGES + 8;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;            // no slack: the window is 1024-aligned
  static constexpr int TMEM_COLS = (2 * MT * BN <= 128) ? 128 : (2 * MT * BN <= 256) ? 256 : 512;
  static_assert(2 * MT * BN <= 512, "TMEM holds two accumulator sets");
  static_assert(!TM || (BN == 128 && MT == 2 && KS == 1), "transposed mode: 128 channels x 256 pixels");
};

struct TileCoord {
  int phase, n0, tx, ty, tb;
};

// ty counts tiles of mt row boxes (p.tiles_y counts 128-row boxes)
__device__ __forceinline__ TileCoord decode_tile(const TapGemmParams& p, int tile, int tiles_m, int tiles_n,
                                                  int bn, int mt_per_tile) {
  TileCoord c;
  const int tpp = tiles_m * tiles_n;
  c.phase = tile / tpp;
  const int r = tile - c.phase * tpp;
  const int nt = r / tiles_m;
  const int mt = r - nt * tiles_m;
  c.n0 = nt * bn;
  const int ty_tiles = p.tiles_y / mt_per_tile;
  c.tx = mt % p.tiles_x;
  c.ty = (mt / p.tiles_x) % ty_tiles;
  c.tb = mt / (p.tiles_x * ty_tiles);
  return c;
}

// Split-K work items: item w of total_tiles * S.  All partial items (split s < S-1) come first,
// the owner items (s = S-1, which reduce and run the epilogue) last.  CTA c walks items c, c + grid,
// ... in increasing order, so on every CTA all partial items precede all owner items: a partial item
// (which never waits) is never queued behind a waiting owner, for any number of rounds.
struct WorkItem {
  int tile, split;
};
__device__ __forceinline__ WorkItem decode_item(int w, int total_tiles, int S) {
  WorkItem it;
  if (S == 1) {
    it.tile = w;
    it.split = 0;
  } else {
    const int npart = total_tiles * (S - 1);
    if (w < npart) {
      it.tile = w / (S - 1);
      it.split = w - it.tile * (S - 1);
    } else {
      it.tile = w - npart;
      it.split = S - 1;
    }
  }
  return it;
}

// Epilogue inner loops, specialised on the 16-bit output format (H = IEEE fp16, else bf16) so that the
// format test stays out of the unrolled loops.
template <bool H>
__device__ __forceinline__ void tm_stage_column(const uint32_t* v, float alpha, float bvs, uint32_t sbuf_addr, int half,
                                                uint32_t chunk, uint32_t col_byte, float& s, float& qq) {
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const uint32_t prow = (uint32_t)(half * 64 + i);          // pixel row within the 128-row box
    const float f = fmaf(__uint_as_float(v[i]), alpha, bvs);
    unsigned short hb;
    float r;
    if (H) {
      const __half hh = __float2half_rn(f);
      r = __half2float(hh);
      hb = __half_as_ushort(hh);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(f);
      r = __bfloat162float(h);
      hb = __bfloat16_as_ushort(h);
    }
    s += r;
    qq = fmaf(r, r, qq);
    const uint32_t dst = sbuf_addr + prow * 128u + (((chunk ^ (prow & 7u)) << 4) | col_byte);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(hb) : "memory");
  }
}

template <bool H>
__device__ __forceinline__ void stage_row32(const uint32_t* v, float alpha, const float* bv, uint32_t row_addr, int half,
                                            int row) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(__uint_as_float(v[g * 8 + i]), alpha, bv[g * 8 + i]);
    const bf16x8 pk = H ? pack8_half(f) : pack8(f);
    const uint32_t dst = row_addr + (uint32_t)(((half * 4 + g) ^ (row & 7)) << 4);   // SWIZZLE_128B
    const uint4 u = *reinterpret_cast<const uint4*>(&pk);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                 : "memory");
  }
}

template <bool H>
__device__ __forceinline__ void column_pair_sums(uint32_t sbuf_addr, int r8, int cp, float& s, float& qq) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r8 * 16 + i;
    const uint32_t a = sbuf_addr + (uint32_t)r * 128u + (uint32_t)((((cp >> 2) ^ (r & 7)) << 4) | ((cp & 3) << 2));
    uint32_t wv;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(a));
    float x0, x1;
    decode2_fmt(wv, H, x0, x1);
    s += x0 + x1;
    qq = fmaf(x0, x0, qq);
    qq = fmaf(x1, x1, qq);
  }
}

// ---- fused GroupNorm(+Swish) epilogue ---------------------------------------------------------------
// The GroupNorm that opens the consumer of this convolution (Model.py:170-173,186-190,132) is applied by the
// epilogue that produces the tensor: pass 1 is the ordinary epilogue (raw 16-bit panel staged, its column sums
// written to the partial-sum array, the raw TMA store skipped when nobody reads the raw tensor); then the tiles
// an image spans meet (gn_peer_rendezvous), every tile reduces the image's partial sums to mean / rstd per
// group, and pass 2 re-reads the fp32 accumulators from TMEM, applies (x - mean) * rstd * gamma + beta (+ Swish)
// and stores the IEEE fp16 result through a second tensor map.  The accumulator buffer is released after pass 2.
//
// Peer tiles = the M tiles of one image with the same N tile: consecutive work items, hence on different,
// co-resident CTAs (grid <= SM count, one CTA per SM).  No deadlock: the arrival of work item w needs only the
// epilogue of item w - grid on the same CTA to finish, which waits for arrivals of items < w - grid + peers <= w,
// so "arrives" is well founded in the item index.  The counter only ever counts up: all `peers` arrivals of one
// launch precede every arrival of the next launch (kernel boundary), so the value a tile's own atomicAdd returns
// tells it which multiple of `peers` to wait for (peers is a power of two, so the grouping survives the 32-bit
// wrap-around) — no reset, no second counter.  Called by one thread between two CTA barriers: its fences are
// cumulative over the partial sums the other threads stored before the first barrier (the pattern of
// cooperative groups' grid sync).
__device__ __forceinline__ void gn_peer_rendezvous(unsigned* counter, unsigned peers) {
  __threadfence();
  const unsigned old = atomicAdd(counter, 1u);
  const unsigned target = (old / peers + 1u) * peers;
  if (old + 1u != target) {
    unsigned cur = old;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(counter) : "memory");
      if ((int)(cur - target) >= 0) break;
      __nanosleep(20);
    }
    if ((int)(cur - target) < 0) __trap();      // a lost peer would otherwise hang the device
  }
  __threadfence();
}

// mean and 1/sqrt(var + eps) of the (image, group) pairs of a tile from the partial sums [img][part][Cout/4] of
// float2 in global memory: T = 256 / pairs (at most 32) consecutive lanes share a pair, each sums every T-th
// entry, xor-shuffle tree, all in double (fixed order; the arithmetic of gn_apply_stats_kernel, groupnorm.cu).
// Every epilogue thread calls it (full-warp shuffles); lane 0 of a pair writes table[pair].
__device__ __forceinline__ void gn_tile_stats(float2* table, const float2* stats, int et, int pairs, int ngt, int g0,
                                              int img0, int n_img, int parts, int nchunk, int cg4, double n, float eps) {
  int T = P_EPI_THREADS / pairs;
  if (T > 32) T = 32;
  const int pr = et / T, l = et - pr * T;
  const int il = pr / ngt, gl = pr - il * ngt;
  const long long bi = img0 + il;
  const bool active = pr < pairs && bi < n_img;
  double s = 0.0, q = 0.0;
  if (active) {
    const float2* base = stats + bi * parts * nchunk + (g0 + gl) * cg4;
    const int E = parts * cg4;
    for (int e = l; e < E; e += T) {
      const int part = e / cg4, kk = e - part * cg4;
      const float2 v = __ldcg(base + (long long)part * nchunk + kk);
      s += (double)v.x;
      q += (double)v.y;
    }
  }
  for (int o = T >> 1; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (l == 0 && pr < pairs) {
    float2 mr = make_float2(0.f, 1.f);
    if (active) {
      const double mean = s / n;
      double var = q / n - mean * mean;    // biased, like nn.GroupNorm
      if (var < 0.0) var = 0.0;
      mr = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    table[pr] = mr;
  }
}

// y = acc * A + B, optional Swish (gn_apply_stats_kernel's default formulation: h + h tanh(h), h = y / 2)
__device__ __forceinline__ float gn_act(float acc, float A, float B, bool silu) {
  const float y = fmaf(acc, A, B);
  return silu ? silu_tanh_half(0.5f * y) : y;
}

// transposed-accumulator variant: one channel (scalar A, B), 64 pixels
__device__ __forceinline__ void tm_stage_column_gn(const uint32_t* v, float A, float B, bool silu, uint32_t sbuf_addr,
                                                   int half, uint32_t chunk, uint32_t col_byte) {
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    const uint32_t prow = (uint32_t)(half * 64 + i);
    const unsigned short hb = __half_as_ushort(__float2half_rn(gn_act(__uint_as_float(v[i]), A, B, silu)));
    const uint32_t dst = sbuf_addr + prow * 128u + (((chunk ^ (prow & 7u)) << 4) | col_byte);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst), "h"(hb) : "memory");
  }
}

// TM ("transposed accumulator", Cout tile of 128 with two row boxes): the MMA takes the WEIGHT tile as
// its A operand (M = 128 channels) and the 256-pixel activation tile as its B operand (N = 256), so a
// k-block is 4 instructions of N = 256 instead of 8 of N = 128 — measured, an N = 128 tcgen05.mma costs
// ~98 clocks against its 64-clock floor while N = 256 runs at its 128-clock floor.  The accumulator is
// then [channel lane][pixel column]; the epilogue transposes it through the swizzled staging panels
// with 16-bit shared-memory stores, and the GroupNorm statistics become per-thread sums.
template <int BN, int STAGES, int MT, int KS, bool TM, bool F16, bool GN>
__global__ void __launch_bounds__(P_THREADS, 1)
tapgemm_persist_kernel(const __grid_constant__ TapGemmParams p, const __grid_constant__ CUtensorMap tmA0,
                       const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                       const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                       const __grid_constant__ CUtensorMap tmGn) {
  using L = PersistSmem<BN, STAGES, MT, KS, TM>;
  // SWIZZLE_128B atoms need 1024-byte alignment; with no static shared memory the dynamic
  // window starts 1024-aligned (checked: a misaligned window traps instead of corrupting)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* staging = smem + L::STAGING_OFF;
  float2* red = reinterpret_cast<float2*>(smem + L::RED_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator buffer complete
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator buffer drained
  uint64_t* pfull_bar = tempty_bar + 2;         // [2] staged output panel complete
  uint64_t* pempty_bar = pfull_bar + 2;         // [2] staged output panel stored
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = p.tiles_x * (p.tiles_y / MT) * p.tiles_b;
  const int tiles_n = (p.Cout + BN - 1) / BN;
  const int total_tiles = tiles_m * tiles_n * p.nphases;
  const int S = p.splits;
  const int total_items = total_tiles * S;

  long long* dbg = nullptr;
  if (p.dbg != nullptr) {
    dbg = p.dbg + (long long)blockIdx.x * 64;
    if (threadIdx.x == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      dbg[0] = (long long)gt;
      dbg[1] = clock64();
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      dbg[6] = smid;
    }
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], P_EPI_THREADS);
      mbar_init(&pfull_bar[b], TM ? P_EPI_THREADS / 2 : P_EPI_THREADS);   // TM: 4 warps fill a panel
      mbar_init(&pempty_bar[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if constexpr (GN) tma_prefetch_desc(&tmGn);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)L::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlaps the tail of the previous kernel under programmatic dependent launch.
  // Nothing below may touch memory the previous kernel writes before griddepcontrol.wait — the
  // producer warp uses the slack to request the WEIGHT tiles of its first ring round (weights are
  // never written on the device), every other thread waits right away.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (dbg != nullptr && threadIdx.x == 0) dbg[2] = clock64();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ----
    // converged warp; one elected lane issues the loads of a stage
    uint32_t npre = 0;   // ring stages whose weight tiles were requested before the wait
    auto produce = [&](const bool pre) {
      uint32_t it = 0;   // k-blocks issued so far (ring position)
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const WorkItem wi = decode_item(w, total_tiles, S);
        const TileCoord c = decode_tile(p, wi.tile, tiles_m, tiles_n, BN, MT);
        const DevPhase& ph = p.phase[c.phase];
        const int kb0 = (ph.nkb * wi.split) / S, kb1 = (ph.nkb * (wi.split + 1)) / S;
        const int wb = (p.w_batch_stride != 0) ? c.tb * p.bb : 0;   // per-image B operand (attention)
        int g = 0;
        for (int t = 0; t < ph.ntaps && g < kb1; ++t) {
          const int si = ph.src[t];
          const DevSrc& s = p.src[si];
          const int ncb = s.C / BK;
          if (g + ncb <= kb0) { g += ncb; continue; }
          const CUtensorMap* tm = (si == 0) ? &tmA0 : (si == 1) ? &tmA1 : &tmA2;
          const int cx = c.tx * p.bw * s.stride + ph.dx[t];
          const int cy = c.ty * (p.bh * MT) * s.stride + ph.dy[t];
          const int cb_img = s.bcast ? 0 : c.tb * p.bb;
          for (int cb = 0; cb < ncb; cb += KS, g += KS) {   // KS = 2: taps and splits hold whole pairs (host)
            if (g < kb0 || g >= kb1) continue;
            if (pre && it >= (uint32_t)STAGES) { npre = it; return; }
            const uint32_t stage = it % STAGES;
            const uint32_t parity = (it / STAGES) & 1u;
            if (!pre) mbar_wait(&empty_bar[stage], parity ^ 1u);
            if (elect_one_sync()) {
              uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
              uint8_t* b_dst = a_dst + KS * MT * A_BYTES;
              if (pre || it >= npre) {
                mbar_expect_tx(&full_bar[stage], (uint32_t)L::STAGE_BYTES);
#pragma unroll
                for (int u = 0; u < KS; ++u)
                  tma_load_3d(b_dst + u * L::B_BYTES, &tmB, &full_bar[stage], ph.w_k0 + (g + u) * BK, c.n0, wb);
              }
              if (!pre) {
#pragma unroll
                for (int u = 0; u < KS; ++u)
                  tma_load_4d(a_dst + u * MT * A_BYTES, tm, &full_bar[stage], (cb + u) * BK, cx, cy, cb_img);
              }
            }
            __syncwarp();
            ++it;
          }
        }
      }
      if (pre) npre = it;
    };
    if (p.w_batch_stride == 0) produce(true);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    produce(false);
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -----
    // the whole warp walks the loop (converged waits), one elected lane issues
    uint32_t it = 0, j = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++j) {
      const WorkItem wi = decode_item(w, total_tiles, S);
      const DevPhase& ph = p.phase[wi.tile / (tiles_m * tiles_n)];
      const int nkb_phase = ph.nkb;
      const int kb0 = (nkb_phase * wi.split) / S;
      const int nkb = (nkb_phase * (wi.split + 1)) / S - kb0;
      const uint32_t idesc_a = make_idesc(BN, ph.fp16_first != 0), idesc_b = make_idesc(BN, ph.fp16_first == 0);
      const int ksw = ph.kb_switch - kb0;   // operand format switches once along K at most
      const uint32_t buf = j & 1u;
      mbar_wait(&tempty_bar[buf], ((j >> 1) & 1u) ^ 1u);   // epilogue has drained this buffer
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + buf * (MT * BN);
      for (int kb = 0; kb < nkb; kb += KS, ++it) {
        const uint32_t stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        const uint32_t idesc = kb < ksw ? idesc_a : idesc_b;   // KS = 2: the switch falls on a pair boundary (host)
        mbar_wait(&full_bar[stage], parity);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + KS * MT * A_BYTES;
          if constexpr (TM) {
            // D[channel][pixel] += W[128 x 64] * Act[256 x 64]^T
            const uint64_t wdesc = make_smem_desc(b_addr), xdesc = make_smem_desc(a_addr);
            const uint32_t idesc_t = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(256 >> 3) << 17);
#pragma unroll
            for (int k2 = 0; k2 < BK / 16; ++k2)
              umma_bf16(d_tmem, wdesc + 2 * k2, xdesc + 2 * k2, idesc_t, (uint32_t)((kb | k2) != 0));
          } else {
#pragma unroll
          for (int u = 0; u < KS; ++u) {
            const uint64_t bdesc = make_smem_desc(b_addr + u * L::B_BYTES);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const uint64_t adesc = make_smem_desc(a_addr + (u * MT + m) * A_BYTES);
#pragma unroll
              for (int k2 = 0; k2 < BK / 16; ++k2)
                umma_bf16(d_tmem + m * BN, adesc + 2 * k2, bdesc + 2 * k2, idesc, (uint32_t)((kb | u | k2) != 0));
            }
          }
          }
          umma_commit(&empty_bar[stage]);
          if (kb + KS >= nkb) umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // --------------------------------------------------- epilogue -----
    // 8 warps: warp w reads TMEM lane quarter w % 4 (rows 32q..32q+31 of the sub-tile) and
    // column half (w - 2) / 4 of the current 64-column panel.  Panels are handed to the
    // store warp through pfull/pempty mbarriers (two staging buffers); the pfull wait doubles
    // as the barrier after which the staged panel may be read back for the statistics.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;              // accumulator row of this thread, 0..127
    const int et = (warp - 2) * 32 + lane;      // epilogue thread id 0..255
    const int rpi = BM / p.bb;                  // rows of one image inside a 128-row sub-tile
    const int rb = row / rpi;                   // image of this thread's row within the sub-tile
    const int nsb_img = 8 / p.bb;               // 16-row statistics sub-blocks per image
    const int tiles_img = p.tiles_x * p.tiles_y;
    const int cp = et & 31, r8 = et >> 5;       // statistics role: column pair, 16-row sub-block
    const int st_ch = et & 15, st_ib = et >> 4; // final reduce role: 4-channel chunk, image in sub-tile
    float2* prev_dst = nullptr;                 // deferred final reduce of the previous panel
    constexpr bool fuse_gn = GN;                // GroupNorm(+Swish) of this tile in a second epilogue pass
    const bool gn_silu = p.gn_silu != 0;
    const int gn_cg = fuse_gn ? p.Cout / p.gn_groups : 4;   // channels per group (multiple of 4)
    const float2* stats2 = reinterpret_cast<const float2*>(p.stats);
    uint32_t j = 0, pc = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++j) {
      const WorkItem wi = decode_item(w, total_tiles, S);
      const TileCoord c = decode_tile(p, wi.tile, tiles_m, tiles_n, BN, MT);
      const uint32_t buf = j & 1u;
      constexpr int NPANEL = BN / PANEL_COLS;
      if (S > 1) {
        int* flags = reinterpret_cast<int*>(p.ws);
        float* parts = p.ws + p.ws_flag_words;
        if (wi.split < S - 1) {
          // partial item: raw fp32 accumulators -> workspace row (128-byte runs per thread), then
          // publish.  No TMA stores, no statistics.
          mbar_wait(&tfull_bar[buf], (j >> 1) & 1u);
          tcgen05_fence_after();
          float* wsrow = parts + ((long long)(wi.tile * (S - 1) + wi.split) * BM + row) * BN + half * 32;
#pragma unroll 1
          for (int pn = 0; pn < NPANEL; ++pn) {
            uint32_t v[32];
            tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) +
                                 (uint32_t)(pn * PANEL_COLS + half * 32), v);
            tmem_wait_ld();
            float4* dst = reinterpret_cast<float4*>(wsrow + pn * PANEL_COLS);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              __stcg(dst + i, make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                          __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
          }
          tcgen05_fence_before();
          mbar_arrive(&tempty_bar[buf]);
          __threadfence();
          named_bar_sync(1, P_EPI_THREADS);
          if (et == 0) {
            int* f = flags + wi.tile * (S - 1) + wi.split;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
          }
          continue;
        }
        // owner item: wait until every partial of this tile has been published
        if (et == 0) {
          for (int sp = 0; sp < S - 1; ++sp) {
            int* f = flags + wi.tile * (S - 1) + sp;
            int seen = 0;
            for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
              if (seen != 0) break;
              __nanosleep(64);
            }
            if (seen == 0) __trap();     // a lost partial would otherwise hang the device
            *f = 0;                      // self-cleaning: the next launch finds the flags zeroed
          }
        }
        named_bar_sync(1, P_EPI_THREADS);
      }
      if constexpr (TM) {
        // thread = (channel lane, 64-pixel quarter of the current row box); two rounds, one per row box
        const int cglob = c.n0 + row;                 // output channel of this thread
        const int bimg = c.tb;                        // bb == 1
        float bvs = p.bias ? __ldg(p.bias + cglob) : 0.f;
        if (p.vec) bvs += __ldg(p.vec + (long long)bimg * p.vec_stride + cglob);
        if (p.vec2) bvs += __ldg(p.vec2 + (long long)bimg * p.vec2_stride + cglob);
        mbar_wait(&tfull_bar[buf], (j >> 1) & 1u);
        tcgen05_fence_after();
        if (dbg != nullptr && et == 0 && j < 8) dbg[8 + j] = clock64();
        const int pn = q >> 1;                        // 64-channel panel this warp's channels belong to
        const uint32_t cc = (uint32_t)((q & 1) * 32 + lane);   // channel within the panel
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          const uint32_t P = pc + 2 * m + pn;         // global panel counter of (row box m, panel pn)
          const uint32_t sb = P & 1u;
          uint8_t* sbuf = staging + sb * PANEL_BYTES;
          uint32_t v[64];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) +
                                 (uint32_t)(m * 128 + half * 64);
          tmem_ld32_nowait(taddr, v);
          tmem_ld32_nowait(taddr + 32, v + 32);
          mbar_wait(&pempty_bar[sb], ((P >> 1) & 1u) ^ 1u);
          tmem_wait_ld();
          if (m == 1 && !fuse_gn) {
            tcgen05_fence_before();
            mbar_arrive(&tempty_bar[buf]);
          }
          float s = 0.f, qq = 0.f;
          const uint32_t col_byte = (cc & 7u) * 2u;
          const uint32_t chunk = cc >> 3;
          tm_stage_column<F16>(v, p.alpha, bvs, smem_u32(sbuf), half, chunk, col_byte, s, qq);
          fence_proxy_async_smem();
          mbar_arrive(&pfull_bar[sb]);
          if (p.stats != nullptr) {
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            qq += __shfl_xor_sync(0xffffffffu, qq, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            qq += __shfl_xor_sync(0xffffffffu, qq, 2);
            if ((lane & 3) == 0) {
              const int ty = c.ty * MT + m;
              const int part = ((c.phase * tiles_img + ty * p.tiles_x + c.tx) << 1) + half;
              reinterpret_cast<float2*>(p.stats)[((long long)bimg * p.stats_parts + part) * (p.Cout >> 2) + (cglob >> 2)] =
                  make_float2(s, qq);
            }
          }
        }
        pc += 4;
        if constexpr (GN) {
          // every partial sum of this tile is on its way to global memory; meet the other tiles of the image
          if (p.gn_peers > 1) {
            named_bar_sync(1, P_EPI_THREADS);
            if (et == 0)
              gn_peer_rendezvous(reinterpret_cast<unsigned*>(p.gn_sync) + c.tb * tiles_n + c.n0 / BN, (unsigned)p.gn_peers);
          } else {
            __threadfence();
          }
          named_bar_sync(1, P_EPI_THREADS);
          const int ngt = BN / gn_cg, g0 = c.n0 / gn_cg;
          gn_tile_stats(red, stats2, et, ngt, ngt, g0, bimg, p.B, p.stats_parts, p.Cout >> 2, gn_cg >> 2,
                        (double)gn_cg * (double)(p.Hm * p.Wm), p.gn_eps);
          named_bar_sync(1, P_EPI_THREADS);
          const float2 mr = red[cglob / gn_cg - g0];
          const float sc = mr.y * __ldg(p.gn_gamma + cglob);
          const float gA = p.alpha * sc;
          const float gB = fmaf(bvs, sc, __ldg(p.gn_beta + cglob) - mr.x * sc);
#pragma unroll 1
          for (int m = 0; m < 2; ++m) {
            const uint32_t P = pc + 2 * m + pn;
            const uint32_t sb = P & 1u;
            uint8_t* sbuf = staging + sb * PANEL_BYTES;
            uint32_t v[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) +
                                   (uint32_t)(m * 128 + half * 64);
            tmem_ld32_nowait(taddr, v);
            tmem_ld32_nowait(taddr + 32, v + 32);
            mbar_wait(&pempty_bar[sb], ((P >> 1) & 1u) ^ 1u);
            tmem_wait_ld();
            if (m == 1) {                      // accumulators fully read (twice): release the buffer
              tcgen05_fence_before();
              mbar_arrive(&tempty_bar[buf]);
            }
            tm_stage_column_gn(v, gA, gB, gn_silu, smem_u32(sbuf), half, cc >> 3, (cc & 7u) * 2u);
            fence_proxy_async_smem();
            mbar_arrive(&pfull_bar[sb]);
          }
          pc += 4;
          named_bar_sync(1, P_EPI_THREADS);    // red[] holds the statistics table until every thread is done
        }
        if (dbg != nullptr && et == 0 && j < 8) dbg[24 + j] = clock64();
        continue;
      }
      int b = c.tb * p.bb + rb;
      if (b >= p.B) b = p.B - 1;         // rows of a ragged last tile: values are never stored
      const float* vrow = p.vec ? p.vec + (long long)b * p.vec_stride : nullptr;
      const float* vrow2 = p.vec2 ? p.vec2 + (long long)b * p.vec2_stride : nullptr;
#pragma unroll 1
      for (int pi = 0; pi < MT * NPANEL; ++pi, ++pc) {
        const int m = pi / NPANEL, pn = pi - m * NPANEL;
        const int n = c.n0 + pn * PANEL_COLS;
        const int nc = n + half * 32;
        const int ty = c.ty * MT + m;    // 128-row box index within the image
        const uint32_t sb = pc & 1u, spar = (pc >> 1) & 1u;
        uint8_t* sbuf = staging + sb * PANEL_BYTES;
        // per-column additive terms first: their latency hides behind the waits below
        float bv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bv[i] = 0.f;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + i);
            bv[4 * i] = t4.x; bv[4 * i + 1] = t4.y; bv[4 * i + 2] = t4.z; bv[4 * i + 3] = t4.w;
          }
        }
        if (vrow) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(vrow + nc) + i);
            bv[4 * i] += t4.x; bv[4 * i + 1] += t4.y; bv[4 * i + 2] += t4.z; bv[4 * i + 3] += t4.w;
          }
        }
        if (vrow2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(vrow2 + nc) + i);
            bv[4 * i] += t4.x; bv[4 * i + 1] += t4.y; bv[4 * i + 2] += t4.z; bv[4 * i + 3] += t4.w;
          }
        }
        if (pi == 0) {
          mbar_wait(&tfull_bar[buf], (j >> 1) & 1u);
          tcgen05_fence_after();
          if (dbg != nullptr && et == 0 && j < 8) dbg[8 + j] = clock64();
        }
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) +
                               (uint32_t)(m * BN + pn * PANEL_COLS + half * 32);
        tmem_ld32_nowait(taddr, v);
        mbar_wait(&pempty_bar[sb], spar ^ 1u);   // the store that last read sbuf has drained it
        tmem_wait_ld();
        if (pi == MT * NPANEL - 1 && !fuse_gn) {   // accumulators fully read: hand the buffer back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(&tempty_bar[buf]);
        }
        if (S > 1) {
          // fixed summation order: partial 0 + partial 1 + ... + own K range
          const float* prow = p.ws + p.ws_flag_words +
                              ((long long)(wi.tile * (S - 1)) * BM + row) * BN + pn * PANEL_COLS + half * 32;
          float acc[32];
          for (int sp = 0; sp < S - 1; ++sp) {
            const float4* src = reinterpret_cast<const float4*>(prow + (long long)sp * BM * BN);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 t4 = __ldcg(src + i);
              if (sp == 0) {
                acc[4 * i] = t4.x; acc[4 * i + 1] = t4.y; acc[4 * i + 2] = t4.z; acc[4 * i + 3] = t4.w;
              } else {
                acc[4 * i] += t4.x; acc[4 * i + 1] += t4.y; acc[4 * i + 2] += t4.z; acc[4 * i + 3] += t4.w;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(acc[i] + __uint_as_float(v[i]));
        }
        const uint32_t row_addr = smem_u32(sbuf) + (uint32_t)row * 128u;
        stage_row32<F16>(v, p.alpha, bv, row_addr, half, row);
        fence_proxy_async_smem();
        mbar_arrive(&pfull_bar[sb]);
        if (p.stats != nullptr && p.bb > 8) {
          // maps smaller than 4x4 (fewer than 16 rows per image: MainCondition.py's 2x2 / 1x1 levels): the
          // 16-row sub-block scheme below does not apply; sum each image's rows directly (tiny work)
          mbar_wait(&pfull_bar[sb], spar);
          for (int idx = et; idx < p.bb * 16; idx += P_EPI_THREADS) {
            const int ib = idx >> 4, ch = idx & 15;
            const int bi = c.tb * p.bb + ib;
            if (bi >= p.B) continue;
            float s = 0.f, qq = 0.f;
            for (int r = ib * rpi; r < (ib + 1) * rpi; ++r) {
              const uint32_t a = smem_u32(sbuf) + (uint32_t)r * 128u +
                                 (uint32_t)((((ch >> 1) ^ (r & 7)) << 4) | ((ch & 1) << 3));
              uint32_t w0, w1;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(a));
              float x0, x1, x2, x3;
              decode2_fmt(w0, F16, x0, x1);
              decode2_fmt(w1, F16, x2, x3);
              s += (x0 + x1) + (x2 + x3);
              qq = fmaf(x0, x0, qq); qq = fmaf(x1, x1, qq); qq = fmaf(x2, x2, qq); qq = fmaf(x3, x3, qq);
            }
            reinterpret_cast<float2*>(p.stats)[((long long)bi * p.stats_parts + c.phase) * (p.Cout >> 2) + (n >> 2) + ch] =
                make_float2(s, qq);
          }
        } else if (p.stats != nullptr) {
          mbar_wait(&pfull_bar[sb], spar);       // every thread's part of the panel is staged
          if (prev_dst != nullptr) {             // final reduce of the previous panel's sub-block sums
            const float2* rd = red + ((pc + 1) & 1u) * 128;
            float S = 0.f, Q = 0.f;
            for (int k = 0; k < nsb_img; ++k) {
              const float2 t2 = rd[(st_ib * nsb_img + k) * 16 + st_ch];
              S += t2.x;
              Q += t2.y;
            }
            *prev_dst = make_float2(S, Q);
          }
          // column sums over the staged bf16 panel: thread = (16-row sub-block, column pair)
          float s = 0.f, qq = 0.f;
          column_pair_sums<F16>(smem_u32(sbuf), r8, cp, s, qq);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          qq += __shfl_xor_sync(0xffffffffu, qq, 1);
          if ((cp & 1) == 0) red[sb * 128 + r8 * 16 + (cp >> 1)] = make_float2(s, qq);
          prev_dst = nullptr;
          const int bi = c.tb * p.bb + st_ib;
          if (et < 16 * p.bb && bi < p.B) {
            const int part = (p.bb == 1) ? c.phase * tiles_img + ty * p.tiles_x + c.tx : c.phase;
            prev_dst = reinterpret_cast<float2*>(p.stats) +
                       ((long long)bi * p.stats_parts + part) * (p.Cout >> 2) + (n >> 2) + st_ch;
          }
        }
      }
      if constexpr (GN) {
        // (1) this tile's partial sums complete in global memory (the last panel's final reduce is still pending)
        named_bar_sync(1, P_EPI_THREADS);
        if (prev_dst != nullptr) {
          const float2* rd = red + ((pc + 1) & 1u) * 128;
          float S1 = 0.f, Q1 = 0.f;
          for (int k = 0; k < nsb_img; ++k) {
            const float2 t2 = rd[(st_ib * nsb_img + k) * 16 + st_ch];
            S1 += t2.x;
            Q1 += t2.y;
          }
          *prev_dst = make_float2(S1, Q1);
          prev_dst = nullptr;
        }
        // (2) meet the other tiles of the image (maps larger than one tile)
        if (p.gn_peers > 1) {
          named_bar_sync(1, P_EPI_THREADS);
          if (et == 0)
            gn_peer_rendezvous(reinterpret_cast<unsigned*>(p.gn_sync) + c.tb * tiles_n + c.n0 / BN, (unsigned)p.gn_peers);
        } else {
          __threadfence();
        }
        named_bar_sync(1, P_EPI_THREADS);
        // (3) mean / rstd of every (image of this tile, group of this N tile)
        const int NG = BN / gn_cg, g0 = c.n0 / gn_cg;      // groups inside this N tile, the first of them
        gn_tile_stats(red, stats2, et, p.bb * NG, NG, g0, c.tb * p.bb, p.B, p.stats_parts, p.Cout >> 2, gn_cg >> 2,
                      (double)gn_cg * (double)(p.Hm * p.Wm), p.gn_eps);
        named_bar_sync(1, P_EPI_THREADS);
        // (4) second pass over the accumulators: normalise, affine, Swish, fp16, staged panel, TMA store.
        // Scale and shift are formed per 4-channel chunk right where they are used (register pressure).
#pragma unroll 1
        for (int pi = 0; pi < MT * NPANEL; ++pi, ++pc) {
          const int m = pi / NPANEL, pn = pi - m * NPANEL;
          const int nc = c.n0 + pn * PANEL_COLS + half * 32;
          const uint32_t sb = pc & 1u, spar = (pc >> 1) & 1u;
          uint8_t* sbuf = staging + sb * PANEL_BYTES;
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (MT * BN) +
                                 (uint32_t)(m * BN + pn * PANEL_COLS + half * 32);
          tmem_ld32_nowait(taddr, v);
          mbar_wait(&pempty_bar[sb], spar ^ 1u);
          tmem_wait_ld();
          if (pi == MT * NPANEL - 1) {       // accumulators read for the second time: release the buffer
            tcgen05_fence_before();
            mbar_arrive(&tempty_bar[buf]);
          }
          const float* prow = p.ws + p.ws_flag_words +
                              ((long long)(wi.tile * (S - 1)) * BM + row) * BN + pn * PANEL_COLS + half * 32;
          const uint32_t row_addr = smem_u32(sbuf) + (uint32_t)row * 128u;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {               // 8 columns = one 16-byte staging store
            float f[8];
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {             // 4-channel chunks never straddle a group
              const int k = 2 * k8 + h4;
              float x[4] = {__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]),
                            __uint_as_float(v[4 * k + 3])};
              if (S > 1) {                               // same order as pass 1: (partial 0 + partial 1 + ...) + own
                float4 acc = __ldcg(reinterpret_cast<const float4*>(prow) + k);
                for (int sp = 1; sp < S - 1; ++sp) {
                  const float4 t4 = __ldcg(reinterpret_cast<const float4*>(prow + (long long)sp * BM * BN) + k);
                  acc.x += t4.x; acc.y += t4.y; acc.z += t4.z; acc.w += t4.w;
                }
                x[0] = acc.x + x[0]; x[1] = acc.y + x[1]; x[2] = acc.z + x[2]; x[3] = acc.w + x[3];
              }
              float add[4] = {0.f, 0.f, 0.f, 0.f};
              if (p.bias) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + k);
                add[0] = t4.x; add[1] = t4.y; add[2] = t4.z; add[3] = t4.w;
              }
              if (vrow) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(vrow + nc) + k);
                add[0] += t4.x; add[1] += t4.y; add[2] += t4.z; add[3] += t4.w;
              }
              if (vrow2) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(vrow2 + nc) + k);
                add[0] += t4.x; add[1] += t4.y; add[2] += t4.z; add[3] += t4.w;
              }
              const float2 mr = red[rb * NG + (nc + 4 * k) / gn_cg - g0];
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + nc) + k);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + nc) + k);
              const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float sc = mr.y * gg[jj];
                // raw value = acc * alpha + add;  y = (raw - mean) * rstd * gamma + beta
                f[4 * h4 + jj] = gn_act(x[jj], p.alpha * sc, fmaf(add[jj], sc, bt[jj] - mr.x * sc), gn_silu);
              }
            }
            const bf16x8 pk = pack8_half(f);
            const uint32_t dst = row_addr + (uint32_t)(((half * 4 + k8) ^ (row & 7)) << 4);   // SWIZZLE_128B
            const uint4 u = *reinterpret_cast<const uint4*>(&pk);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                         : "memory");
          }
          fence_proxy_async_smem();
          mbar_arrive(&pfull_bar[sb]);
        }
        named_bar_sync(1, P_EPI_THREADS);      // red[] (the statistics table) is reused by the next tile's pass 1
      }
      if (dbg != nullptr && et == 0 && j < 8) dbg[24 + j] = clock64();
    }
    if (p.stats != nullptr) {                    // flush the last panel's statistics
      named_bar_sync(1, P_EPI_THREADS);
      if (prev_dst != nullptr) {
        const float2* rd = red + ((pc + 1) & 1u) * 128;
        float S = 0.f, Q = 0.f;
        for (int k = 0; k < nsb_img; ++k) {
          const float2 t2 = rd[(st_ib * nsb_img + k) * 16 + st_ch];
          S += t2.x;
          Q += t2.y;
        }
        *prev_dst = make_float2(S, Q);
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------ output stores ----
    if (lane == 0) {
      uint32_t pc = 0;
      constexpr int NPANEL = BN / PANEL_COLS;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const WorkItem wi = decode_item(w, total_tiles, S);
        if (wi.split < S - 1) continue;          // partial items store nothing
        const TileCoord c = decode_tile(p, wi.tile, tiles_m, tiles_n, BN, MT);
        const DevPhase& ph = p.phase[c.phase];
        // pass 0: the raw tensor (skipped when only its GroupNorm is wanted); pass 1: the fused GroupNorm output
        constexpr int npass = GN ? 2 : 1;
        for (int pass = 0; pass < npass; ++pass) {
          for (int pi = 0; pi < MT * NPANEL; ++pi, ++pc) {
            const int m = pi / NPANEL, pn = pi - m * NPANEL;
            const int n = c.n0 + pn * PANEL_COLS;
            const int ty = c.ty * MT + m;
            const uint32_t sb = pc & 1u, spar = (pc >> 1) & 1u;
            const uint8_t* sbuf = staging + sb * PANEL_BYTES;
            mbar_wait(&pfull_bar[sb], spar);
            if (pass == 0 && p.gn_only) {        // staged only for the statistics
              mbar_arrive(&pempty_bar[sb]);
              continue;
            }
            if (pass == 1)
              tma_store_3d(&tmGn, sbuf, n, c.tx * p.bw, c.tb * p.bb * p.Hm + ty * p.bh);
            else if (p.out_scale == 1)
              tma_store_3d(&tmOut, sbuf, n, c.tx * p.bw, c.tb * p.bb * p.Hm + ty * p.bh);
            else
              tma_store_5d(&tmOut, sbuf, n, ph.px, c.tx * p.bw, ph.py, c.tb * p.bb * p.Hm + ty * p.bh);
            bulk_commit_group();
            bulk_wait_group_read<0>();           // shared memory has been read: the buffer is free again
            mbar_arrive(&pempty_bar[sb]);
          }
        }
      }
      bulk_wait_group<0>();                      // all tensor stores complete before the CTA retires
    }
  }

  if (dbg != nullptr && warp == 2 && lane == 0) dbg[3] = clock64();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)L::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------- host ----
int device_sm_count() {
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& c = n[dev & 63];
  if (c == 0 && (cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c <= 0)) c = 148;
  return c;
}

static int persist_bn(const its_conv_desc* d, const TapGemmParams& p) {
  if (d->bn != 0) return d->bn;
  return (p.Cout % 256 == 0) ? 256 : (p.Cout % 192 == 0) ? 192 : (p.Cout % 128 == 0) ? 128 : 64;
}

struct PersistCfg {
  int bn, mt;
  bool tm;      // transposed accumulator (128 channels x 256 pixels)
};

// Tile configuration of a layer on the persistent schedule.  d->cluster doubles as a tuning knob:
// 1 = one row box per tile, 2 = two row boxes without the transposed accumulator.
static PersistCfg persist_cfg(const its_conv_desc* d, const TapGemmParams& p) {
  PersistCfg c;
  c.bn = persist_bn(d, p);
  // two row boxes per tile (shared weight tiles) when the N tile leaves TMEM room for two
  // double-buffered accumulators and a tile of 2 x bh rows still tiles the image
  c.mt = (c.bn <= 128 && p.bb == 1 && p.tiles_y % 2 == 0 && d->cluster != 1 && p.splits == 1) ? 2 : 1;
  c.tm = (c.mt == 2 && c.bn == 128 && d->cluster != 2);
  return c;
}

int tapgemm_stats_parts(const its_conv_desc* d, const TapGemmParams& p) {
  // one slot per 128-row box of the image (two 64-pixel halves per box in transposed mode)
  const PersistCfg c = persist_cfg(d, p);
  return p.nphases * (p.bb == 1 ? p.tiles_x * p.tiles_y * (c.tm ? 2 : 1) : 1);
}

int tapgemm_gn_sync_words(const its_conv_desc* d, const TapGemmParams& p) {
  if (p.nphases != 1 || p.out_scale != 1 || p.w_batch_stride != 0 || p.bb > 8) return -1;
  if (p.gn_groups <= 0 || p.Cout % p.gn_groups != 0) return -1;
  const int cg = p.Cout / p.gn_groups;
  const PersistCfg c = persist_cfg(d, p);
  // a group lies inside one N tile and a 4-channel statistics chunk inside one group; the (image, group) pairs of
  // a tile share the 256 epilogue threads evenly (powers of two)
  if (cg % 4 != 0 || c.bn % cg != 0) return -1;
  const int pairs = p.bb * (c.bn / cg);
  if (pairs > P_EPI_THREADS || (pairs & (pairs - 1)) != 0) return -1;
  const int peers = p.tiles_x * (p.tiles_y / c.mt);
  if (peers == 1) return 0;
  if (peers > 64 || (peers & (peers - 1)) != 0) return -1;   // power of two: see gn_peer_rendezvous
  return p.tiles_b * (p.Cout / c.bn);
}

bool tapgemm_persist_eligible(const its_conv_desc* d, const TapGemmParams& p) {
  if (p.out_fp32 || p.out_nchw || p.res != nullptr) return false;
  if (p.w_batch_stride != 0 && (p.bb != 1 || p.splits > 1)) return false;
  if (p.Cout % 64 != 0 || p.out_c_pitch % 8 != 0) return false;
  const int bn = persist_bn(d, p);
  if (!(bn == 64 || bn == 128 || bn == 192 || bn == 256) || p.Cout % bn != 0) return false;
  if (p.bw * p.bh * p.bb != BM || p.Wm % p.bw != 0 || p.Hm % p.bh != 0) return false;
  if (p.bb > 1 && p.bh != p.Hm) return false;
  return true;
}

template <int BN, int STAGES, int MT, int KS, bool TM, bool F16, bool GN>
static int launch_persist_fmt(const TapGemmParams& p, const CUtensorMap* tmA, const CUtensorMap& tmB,
                              const CUtensorMap& tmOut, const CUtensorMap& tmGn, cudaStream_t stream) {
  using L = PersistSmem<BN, STAGES, MT, KS, TM>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  auto kern = tapgemm_persist_kernel<BN, STAGES, MT, KS, TM, F16, GN>;
  static PerDeviceBytes configured;
  if (configured.need(L::TOTAL))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int tiles = p.tiles_x * (p.tiles_y / MT) * p.tiles_b * (p.Cout / BN) * p.nphases * p.splits;
  const int sms = device_sm_count();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles < sms ? tiles : sms, 1, 1);
  cfg.blockDim = dim3(P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(1) ? 1 : 0;
  ITS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p, tmA[0], tmA[1], tmA[2], tmB, tmOut, tmGn));
  return ITS_OK;
}

// the 16-bit output format (bf16 / IEEE fp16) is a compile-time parameter of the kernel: a run-time test inside
// the unrolled epilogue loops cost 1.3-1.8 % of a config-A pass (measured)
template <int BN, int STAGES, int MT, int KS, bool TM = false>
static int launch_persist(const TapGemmParams& p, const CUtensorMap* tmA, const CUtensorMap& tmB,
                          const CUtensorMap& tmOut, const CUtensorMap& tmGn, cudaStream_t stream) {
  // ... and so is the fused GroupNorm pass: the plain epilogue keeps its register allocation
  if (p.gn_out != nullptr) {
    if (p.out_fp16) return launch_persist_fmt<BN, STAGES, MT, KS, TM, true, true>(p, tmA, tmB, tmOut, tmGn, stream);
    return launch_persist_fmt<BN, STAGES, MT, KS, TM, false, true>(p, tmA, tmB, tmOut, tmGn, stream);
  }
  if (p.out_fp16) return launch_persist_fmt<BN, STAGES, MT, KS, TM, true, false>(p, tmA, tmB, tmOut, tmGn, stream);
  return launch_persist_fmt<BN, STAGES, MT, KS, TM, false, false>(p, tmA, tmB, tmOut, tmGn, stream);
}

int tapgemm_launch_persist(const its_conv_desc* d, const TapGemmParams& p, cudaStream_t stream) {
  const int bn = persist_bn(d, p);
  ITS_REQUIRE(tapgemm_persist_eligible(d, p), "its_conv_igemm: layer not eligible for the persistent kernel");
  ITS_REQUIRE(p.w_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0, "its_conv_igemm: weight alignment");
  ITS_REQUIRE((reinterpret_cast<uintptr_t>(p.out) & 15) == 0, "its_conv_igemm: output pointer alignment");
  if (p.stats != nullptr)
    ITS_REQUIRE(p.stats_parts == tapgemm_stats_parts(d, p), "its_conv_igemm: stats_parts=%d, the tiling writes %d",
                p.stats_parts, tapgemm_stats_parts(d, p));
  const PersistCfg cfg = persist_cfg(d, p);
  const int mt = cfg.mt;
  TapGemmParams pp = p;
  if (p.splits > 1) {
    // workspace: [flags: one int per (tile, partial split), padded to 64 words][partials fp32]
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_b * (p.Cout / bn) * p.nphases;
    const long long nflag = ((tiles * (p.splits - 1) + 63) / 64) * 64;
    const long long need = nflag + tiles * (p.splits - 1) * BM * bn;
    ITS_REQUIRE(d->ws != nullptr && d->ws_elems >= need,
                "its_conv_igemm: persistent split-K workspace has %lld floats, %lld needed (must be zero-initialised)",
                (long long)d->ws_elems, need);
    for (int f = 0; f < p.nphases; ++f)
      ITS_REQUIRE(p.splits <= p.phase[f].nkb, "its_conv_igemm: splits=%d exceeds the %d k-blocks of phase %d", p.splits, p.phase[f].nkb, f);
    // No limit on tiles * splits: the static round-robin schedule hands CTA c the items c, c + grid,
    // c + 2 grid, ... in increasing order and every partial item precedes every owner item in that
    // order, so no partial item is ever queued behind a (possibly waiting) owner on any CTA; partial
    // items never wait, hence every owner's spin terminates whatever the number of rounds.
    pp.ws_flag_words = (int)nflag;
  }
  if (p.gn_out != nullptr) {
    const int words = tapgemm_gn_sync_words(d, p);
    ITS_REQUIRE(words >= 0, "its_conv_igemm: this layer's GroupNorm cannot be fused into its epilogue "
                "(its_conv_gn_sync_words() < 0): Cout=%d groups=%d bn=%d", p.Cout, p.gn_groups, bn);
    ITS_REQUIRE(p.stats != nullptr, "its_conv_igemm: the fused GroupNorm epilogue needs the statistics array");
    ITS_REQUIRE(p.gn_gamma != nullptr && p.gn_beta != nullptr, "its_conv_igemm: gn_gamma / gn_beta");
    ITS_REQUIRE(words == 0 || p.gn_sync != nullptr, "its_conv_igemm: gn_sync needs %d zero-initialised ints", words);
    ITS_REQUIRE(p.gn_c_pitch >= p.Cout && p.gn_c_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(p.gn_out) & 15) == 0,
                "its_conv_igemm: gn_out pitch / alignment");
    pp.gn_sync_words = words;
    pp.gn_peers = p.tiles_x * (p.tiles_y / mt);
  }
  CUtensorMap tmA[ITS_MAX_SRC], tmB, tmOut, tmGn;
  int rc = tapgemm_encode_operand_maps(p, bn, tmA, &tmB, mt);
  if (rc != ITS_OK) return rc;
  const cuuint64_t pitch_b = (cuuint64_t)p.out_c_pitch * 2;
  if (p.out_scale == 1) {
    // [B*Hm][Wm][C]: image and row merge into one dimension (a tile spanning images covers whole images)
    const cuuint64_t dims[3] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wm, (cuuint64_t)p.B * p.Hm};
    const cuuint64_t strides[2] = {pitch_b, pitch_b * p.Wout};
    const cuuint32_t box[3] = {(cuuint32_t)PANEL_COLS, (cuuint32_t)p.bw, (cuuint32_t)(p.bh * p.bb)};
    const cuuint32_t estr[3] = {1, 1, 1};
    rc = encode_bf16_map(&tmOut, 3, p.out, dims, strides, box, estr, "output");
  } else {
    // sub-pixel phases: [B*Hm][py][Wm][px][C] view of the [B][2Hm][2Wm][C] tensor
    const cuuint64_t dims[5] = {(cuuint64_t)p.Cout, (cuuint64_t)p.out_scale, (cuuint64_t)p.Wm,
                                (cuuint64_t)p.out_scale, (cuuint64_t)p.B * p.Hm};
    const cuuint64_t strides[4] = {pitch_b, pitch_b * p.out_scale, pitch_b * p.Wout,
                                   pitch_b * p.Wout * p.out_scale};
    const cuuint32_t box[5] = {(cuuint32_t)PANEL_COLS, 1, (cuuint32_t)p.bw, 1, (cuuint32_t)(p.bh * p.bb)};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    rc = encode_bf16_map(&tmOut, 5, p.out, dims, strides, box, estr, "output");
  }
  if (rc != ITS_OK) return rc;
  tmGn = tmOut;
  if (p.gn_out != nullptr) {
    const cuuint64_t gpitch = (cuuint64_t)p.gn_c_pitch * 2;
    const cuuint64_t dims[3] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wm, (cuuint64_t)p.B * p.Hm};
    const cuuint64_t strides[2] = {gpitch, gpitch * p.Wm};
    const cuuint32_t box[3] = {(cuuint32_t)PANEL_COLS, (cuuint32_t)p.bw, (cuuint32_t)(p.bh * p.bb)};
    const cuuint32_t estr[3] = {1, 1, 1};
    rc = encode_bf16_map(&tmGn, 3, p.gn_out, dims, strides, box, estr, "fused GroupNorm output");
    if (rc != ITS_OK) return rc;
  }
  // two k-blocks per ring stage when every tap, split and format switch covers whole pairs
  bool pairs = (mt == 1 && bn <= 128 && d->cluster != 3);
  for (int s = 0; s < p.nsrc && pairs; ++s) pairs = (p.src[s].C % (2 * BK) == 0);
  for (int f = 0; f < p.nphases && pairs; ++f)
    pairs = (p.phase[f].nkb % (2 * p.splits) == 0) && (p.phase[f].kb_switch % 2 == 0);
  if (mt == 2) {
    if (bn == 64) return launch_persist<64, 4, 2, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
    if (cfg.tm) return launch_persist<128, 4, 2, 1, true>(pp, tmA, tmB, tmOut, tmGn, stream);
    return launch_persist<128, 4, 2, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
  }
  if (pairs) {
    if (bn == 64) return launch_persist<64, 4, 1, 2>(pp, tmA, tmB, tmOut, tmGn, stream);
    return launch_persist<128, 3, 1, 2>(pp, tmA, tmB, tmOut, tmGn, stream);
  }
  switch (bn) {
    case 64:  return launch_persist<64, 8, 1, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
    case 128: return launch_persist<128, 6, 1, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
    case 192: return launch_persist<192, 4, 1, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
    default:  return launch_persist<256, 4, 1, 1>(pp, tmA, tmB, tmOut, tmGn, stream);
  }
}

}  // namespace its
