// Implicit-GEMM convolution ("tap-GEMM") on Blackwell tensor cores.
//
//   D[128 rows x BN cols] += A_tap[128 x 64] * W[BN x 64]^T     per k-block
//
// * GEMM rows are output pixels (b, y, x); a 128-row tile is a box of
//   bb images x bh rows x bw columns of the NHWC bf16 activation tensor.  For
//   each filter tap the A tile is that same box shifted by (dy, dx): a single
//   4-D tiled TMA load (channels innermost, 64 channels = one 128-byte swizzle
//   row per pixel).  Zero padding is TMA out-of-bounds fill, stride-2
//   convolutions use the tensor map's element strides, nearest-upsample and
//   transposed convolutions run as up to four sub-pixel phases (blockIdx.z) with
//   their own tap tables and weight columns.
// * Weights are a K-major bf16 matrix [Cout][K], loaded by a 3-D TMA (the third
//   dimension is the per-image batch of the attention bmm's).
// * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) is issued by one thread, the
//   accumulator lives in TMEM; four epilogue warps read it back with tcgen05.ld
//   and fuse alpha, bias, the per-image time/label-embedding vector, the
//   residual, and the bf16 (or fp32) store.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//   warps 2..5 = epilogue.  smem ring of STAGES {A 16 KB, B BN*128 B} buffers
//   guarded by full/empty mbarriers; tcgen05.commit releases the slots.
//
// Reference semantics: see its_conv_igemm in include/its_b200.h.
#include "tapgemm.cuh"
#include "sm100_ptx.cuh"

namespace its {

// alpha / bias / per-image vectors / residual, then the store in one of the three
// output formats.  f[8] holds raw accumulators of columns n..n+7 of GEMM row (b,y,x).
__device__ __forceinline__ void apply_and_store8(const TapGemmParams& p, const DevPhase& ph, int b, int y,
                                                 int x, int n, float* f) {
  const int yo = y * p.out_scale + ph.py, xo = x * p.out_scale + ph.px;
  if (p.out_nchw) {  // thin outputs (the 3-channel tail): element-wise guards, coalesced over pixels
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (n + j < p.Cout) {
        float v = f[j] * p.alpha;
        if (p.bias) v += __ldg(p.bias + n + j);
        static_cast<float*>(p.out)[(((long long)b * p.Cout + n + j) * p.Hout + yo) * p.Wout + xo] = v;
      }
    }
    return;
  }
  const long long opix = ((long long)b * p.Hout + yo) * p.Wout + xo;
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] *= p.alpha;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += __ldg(p.bias + n + j);
  }
  if (p.vec) {
    const float* v = p.vec + (long long)b * p.vec_stride + n;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += __ldg(v + j);
  }
  if (p.vec2) {
    const float* v = p.vec2 + (long long)b * p.vec2_stride + n;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += __ldg(v + j);
  }
  if (p.res) {
    float r[8];
    unpack8_fmt(*reinterpret_cast<const bf16x8*>(p.res + opix * p.res_c_pitch + n), r, p.res_fp16 != 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += r[j];
  }
  if (p.out_fp32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + opix * p.out_c_pitch + n);
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    *reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(p.out) + opix * p.out_c_pitch + n) = pack8_fmt(f, p.out_fp16 != 0);
  }
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 8 + 1024;  // + alignment slack
};

// CS > 1: the CS CTAs of a cluster are consecutive M tiles of the same (N tile, phase,
// split).  They need the same weight tile, so each loads 1/CS of it and multicasts the
// slice into all CS shared memories: L2->SM weight traffic drops by CS.  A stage may only
// be refilled once every CTA of the cluster has consumed it, so the MMA issuers commit to
// the "empty" barriers of all CS CTAs.
template <int BN, int STAGES, int MIN_CTAS, int CS>
__global__ void __launch_bounds__(NUM_THREADS, MIN_CTAS)
tapgemm_sm100_kernel(const __grid_constant__ TapGemmParams p, const __grid_constant__ CUtensorMap tmA0,
                     const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ CUtensorMap tmB) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment in the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // diagnostics: [0] globaltimer at entry, [1] clock at entry, [2] after setup, [3] producer done
  // issuing, [4] accumulator complete, [5] epilogue done, [8+kb] clock when k-block kb became full
  long long* dbg = nullptr;
  if (p.dbg != nullptr) {
    const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    dbg = p.dbg + cta * 64;
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
  const int phase_idx = blockIdx.z / p.splits, split = blockIdx.z - phase_idx * p.splits;
  const DevPhase& ph = p.phase[phase_idx];
  // split-K: this CTA owns k-blocks [kb0, kb1) of the phase
  const int kb0 = (ph.nkb * split) / p.splits, kb1 = (ph.nkb * (split + 1)) / p.splits;
  const int nkb = kb1 - kb0;
  const int mt = blockIdx.x;
  const int tx = mt % p.tiles_x;
  const int ty = (mt / p.tiles_x) % p.tiles_y;
  const int tb = mt / (p.tiles_x * p.tiles_y);
  const int n0 = blockIdx.y * BN;
  constexpr int TMEM_COLS = tmem_cols_for(BN);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CS);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers' barriers are initialised before anything remote arrives
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();   // set-up above overlaps the previous kernel's tail; global memory only from here on
  if (dbg != nullptr && threadIdx.x == 0) dbg[2] = clock64();
  const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ----
    if (lane == 0) {
      const int wb = (p.w_batch_stride != 0) ? tb * p.bb : 0;
      int g = 0;  // k-block index within the phase
      for (int t = 0; t < ph.ntaps && g < kb1; ++t) {
        const int si = ph.src[t];
        const DevSrc& s = p.src[si];
        const int ncb = s.C / BK;
        if (g + ncb <= kb0) { g += ncb; continue; }
        const CUtensorMap* tm = (si == 0) ? &tmA0 : (si == 1) ? &tmA1 : &tmA2;
        const int cx = tx * p.bw * s.stride + ph.dx[t];
        const int cy = ty * p.bh * s.stride + ph.dy[t];
        const int cb_img = s.bcast ? 0 : tb * p.bb;
        for (int cb = 0; cb < ncb; ++cb, ++g) {
          if (g < kb0 || g >= kb1) continue;
          const int kb = g - kb0;
          const int stage = kb % STAGES;
          const uint32_t parity = (uint32_t)((kb / STAGES) & 1);
          mbar_wait(&empty_bar[stage], parity ^ 1u);
          uint8_t* a_dst = smem + stage * L::STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)L::STAGE_BYTES);
          tma_load_4d(a_dst, tm, &full_bar[stage], cb * BK, cx, cy, cb_img);
          if (CS == 1) {
            tma_load_3d(b_dst, &tmB, &full_bar[stage], ph.w_k0 + g * BK, n0, wb);
          } else {
            constexpr int SLICE = BN / CS;   // rows of the weight tile this CTA fetches for everyone
            tma_load_3d_mcast(b_dst + crank * (SLICE * BK * 2), &tmB, &full_bar[stage], ph.w_k0 + g * BK,
                              n0 + (int)crank * SLICE, wb, kMask);
          }
        }
      }
      if (dbg != nullptr) dbg[3] = clock64();
    }
  } else if (warp == 1) {
    // ------------------------------------------------- MMA issuer -----
    if (lane == 0) {
      const uint32_t idesc_a = make_idesc(BN, ph.fp16_first != 0), idesc_b = make_idesc(BN, ph.fp16_first == 0);
      const int ksw = ph.kb_switch - kb0;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t idesc = kb < ksw ? idesc_a : idesc_b;
        const int stage = kb % STAGES;
        const uint32_t parity = (uint32_t)((kb / STAGES) & 1);
        mbar_wait(&full_bar[stage], parity);
        tcgen05_fence_after();
        if (dbg != nullptr && kb < 56) dbg[8 + kb] = clock64();
        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(a_addr);
        const uint64_t bdesc = make_smem_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kb | k) != 0));
        if (CS == 1) umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
        else umma_commit_mcast(&empty_bar[stage], kMask);
      }
      umma_commit(tmem_full_bar);        // accumulator complete
    }
    __syncwarp();
  } else {
    // --------------------------------------------------- epilogue -----
    // TMEM gives each thread one accumulator ROW; storing rows straight to global would
    // make every warp store touch 32 different lines (measured: ~4.5 clk per 16-byte
    // request, a 10 us epilogue).  So: (A) dump the fp32 tile into the now idle pipeline
    // buffers, (B) re-read it with lanes running along the channel dimension and do
    // alpha/bias/vectors/residual + the store fully coalesced.
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int et = q * 32 + lane;        // epilogue thread id 0..127 (= accumulator row in phase A)
    constexpr int PITCH = BN + 4;        // fp32 words per staged row (+4: conflict-free float4 access)
    float* stage_f = reinterpret_cast<float*>(smem);
    mbar_wait(tmem_full_bar, 0);         // all MMAs done => every TMA load has landed and been consumed
    tcgen05_fence_after();
    if (dbg != nullptr && warp == 2 && lane == 0) dbg[4] = clock64();
    if (p.out_nchw) {
      // thin output (3-channel tail): lanes are consecutive pixels, already coalesced
      const int rb = et / (p.bh * p.bw), ry = (et / p.bw) % p.bh, rx = et % p.bw;
      const int b = tb * p.bb + rb, y = ty * p.bh + ry, x = tx * p.bw + rx;
      const bool valid = (b < p.B) && (y < p.Hm) && (x < p.Wm);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int n = n0 + c0 + g * 8;
            if (n < p.Cout) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
              apply_and_store8(p, ph, b, y, x, n, f);
            }
          }
        }
      }
    } else {
      // (A) accumulator rows -> shared memory
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        float4* dst = reinterpret_cast<float4*>(stage_f + et * PITCH + c0);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                               __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      // (B) coalesced pass: consecutive threads take consecutive 8-column groups of a row
      constexpr int NCG = BN / 8;
      const long long rows_total = (long long)p.B * p.Hm * p.Wm;
#pragma unroll 1
      for (int i = et; i < BM * NCG; i += 128) {
        const int r = i / NCG, cgi = i - r * NCG;
        const int n = n0 + cgi * 8;
        const int rb = r / (p.bh * p.bw), ry = (r / p.bw) % p.bh, rx = r % p.bw;
        const int b = tb * p.bb + rb, y = ty * p.bh + ry, x = tx * p.bw + rx;
        if (b >= p.B || y >= p.Hm || x >= p.Wm || n >= p.Cout) continue;
        const float4* src = reinterpret_cast<const float4*>(stage_f + r * PITCH + cgi * 8);
        const float4 lo = src[0], hi = src[1];
        float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        if (p.splits > 1) {   // split-K partial: raw accumulators, reduced by the finalize launch
          const long long grow = ((long long)b * p.Hm + y) * p.Wm + x;
          float4* dst = reinterpret_cast<float4*>(p.ws + ((long long)blockIdx.z * rows_total + grow) * p.Cout + n);
          dst[0] = lo;
          dst[1] = hi;
        } else {
          apply_and_store8(p, ph, b, y, x, n, f);
        }
      }
    }
  }

  if (dbg != nullptr && warp == 2 && lane == 0) dbg[5] = clock64();
  tcgen05_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // nobody exits while a peer may still signal or multicast into it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// Split-K second pass: sum the `splits` partial tiles in a fixed order, then the
// same epilogue.  One thread per (phase, GEMM row, 8 output columns).
__global__ void __launch_bounds__(256) tapgemm_finalize_kernel(const TapGemmParams p) {
  pdl_prologue();
  const int nv = p.Cout / 8;
  const long long rows = (long long)p.B * p.Hm * p.Wm;
  const long long total = rows * nv * p.nphases;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % nv);
    const long long r = (i / nv) % rows;
    const int phase_idx = (int)(i / (nv * rows));
    const DevPhase& ph = p.phase[phase_idx];
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float4* src = reinterpret_cast<const float4*>(
          p.ws + (((long long)(phase_idx * p.splits + s) * rows) + r) * p.Cout + cv * 8);
      const float4 a = src[0], b4 = src[1];
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
      f[4] += b4.x; f[5] += b4.y; f[6] += b4.z; f[7] += b4.w;
    }
    const int x = (int)(r % p.Wm);
    const int y = (int)((r / p.Wm) % p.Hm);
    const int b = (int)(r / ((long long)p.Wm * p.Hm));
    apply_and_store8(p, ph, b, y, x, cv * 8, f);
  }
}

// ---------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int encode_bf16_map(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims,
                           const cuuint64_t* strides_bytes, const cuuint32_t* box,
                           const cuuint32_t* estr, const char* what) {
  EncodeTiledFn enc = get_encoder();
  ITS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(ITS_ERR_CUDA,
                     "cuTensorMapEncodeTiled(%s) failed: CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)",
                     what, (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                     (unsigned long long)dims[2], box[0], box[1], box[2]);
  return ITS_OK;
}

// Tensor maps of the A sources (4-D boxes of bb x bh x bw pixels x 64 channels) and of the
// packed weights (3-D: K, Cout, per-image batch; box 64 x b_box_rows x 1).
int tapgemm_encode_operand_maps(const TapGemmParams& p, int b_box_rows, CUtensorMap* tmA, CUtensorMap* tmB_out,
                                int row_boxes) {
  memset(tmA, 0, sizeof(CUtensorMap) * ITS_MAX_SRC);
  for (int s = 0; s < p.nsrc; ++s) {
    const DevSrc& in = p.src[s];
    ITS_REQUIRE((reinterpret_cast<uintptr_t>(in.ptr) & 15) == 0, "its_conv_igemm: src %d pointer alignment", s);
    ITS_REQUIRE(p.bw * in.stride <= 256 && p.bh * row_boxes * in.stride <= 256, "its_conv_igemm: box too large");
    const cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H,
                                (cuuint64_t)(in.bcast ? 1 : p.B)};
    const cuuint64_t strides[3] = {(cuuint64_t)in.c_pitch * 2, (cuuint64_t)in.W * in.c_pitch * 2,
                                   (cuuint64_t)in.H * in.W * in.c_pitch * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(p.bw * in.stride), (cuuint32_t)(p.bh * row_boxes * in.stride),
                               (cuuint32_t)p.bb};
    const cuuint32_t estr[4] = {1, (cuuint32_t)in.stride, (cuuint32_t)in.stride, 1};
    int rc = encode_bf16_map(&tmA[s], 4, in.ptr, dims, strides, box, estr, "activations");
    if (rc != ITS_OK) return rc;
  }
  for (int s = p.nsrc; s < ITS_MAX_SRC; ++s) tmA[s] = tmA[0];

  int k_extent = 0;
  for (int f = 0; f < p.nphases; ++f) {
    const int k1 = p.phase[f].w_k0 + p.phase[f].nkb * BK;
    if (k1 > k_extent) k_extent = k1;
  }
  CUtensorMap& tmB = *tmB_out;
  {
    const long long nbatch = (p.w_batch_stride != 0) ? p.B : 1;
    const cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)p.Cout, (cuuint64_t)nbatch};
    const cuuint64_t strides[2] = {(cuuint64_t)p.w_pitch * 2,
                                   (cuuint64_t)((p.w_batch_stride != 0) ? p.w_batch_stride
                                                                        : (long long)p.Cout * p.w_pitch) * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)b_box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    int rc = encode_bf16_map(&tmB, 3, p.w, dims, strides, box, estr, "weights");
    if (rc != ITS_OK) return rc;
  }
  return ITS_OK;
}

template <int BN, int STAGES, int MIN_CTAS, int CS>
static int launch_variant(const TapGemmParams& p, const CUtensorMap* tmA, const CUtensorMap& tmB,
                          cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = tapgemm_sm100_kernel<BN, STAGES, MIN_CTAS, CS>;
  static PerDeviceBytes configured;
  if (configured.need(L::TOTAL))
    ITS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.tiles_x * p.tiles_y * p.tiles_b, (p.Cout + BN - 1) / BN, p.nphases * p.splits);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(1) ? 2 : 1;
  ITS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p, tmA[0], tmA[1], tmA[2], tmB));
  if (p.splits > 1) {
    const long long items = (long long)p.B * p.Hm * p.Wm * (p.Cout / 8) * p.nphases;
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ITS_LAUNCH(tapgemm_finalize_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, p);
    ITS_CHECK_LAUNCH();
  }
  return ITS_OK;
}

template <int BN, int STAGES, int MIN_CTAS>
static int launch_cs(int cs, const TapGemmParams& p, const CUtensorMap* tmA, const CUtensorMap& tmB,
                     cudaStream_t stream) {
  if (cs == 4) return launch_variant<BN, STAGES, MIN_CTAS, 4>(p, tmA, tmB, stream);
  if (cs == 2) return launch_variant<BN, STAGES, MIN_CTAS, 2>(p, tmA, tmB, stream);
  return launch_variant<BN, STAGES, MIN_CTAS, 1>(p, tmA, tmB, stream);
}

int tapgemm_launch_sm100(const its_conv_desc* d, const TapGemmParams& p, cudaStream_t stream) {
  // tile N
  int bn = d->bn;
  if (bn == 0)
    bn = (p.Cout % 256 == 0) ? 256 : (p.Cout % 192 == 0) ? 192 : (p.Cout % 128 == 0) ? 128 : (p.Cout > 32) ? 64 : 32;
  ITS_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 192 || bn == 256, "its_conv_igemm: bn=%d", bn);
  ITS_REQUIRE(p.out_nchw || p.Cout % 8 == 0, "its_conv_igemm: Cout=%d must be a multiple of 8", p.Cout);
  ITS_REQUIRE(p.splits == 1 || (!p.out_nchw && p.Cout % 8 == 0), "its_conv_igemm: split-K needs NHWC output");
  if (p.splits > 1) {
    const long long need = (long long)p.nphases * p.splits * p.B * p.Hm * p.Wm * p.Cout;
    ITS_REQUIRE(d->ws_elems >= need, "its_conv_igemm: workspace has %lld floats, %lld needed", (long long)d->ws_elems, need);
  }
  for (int f = 0; f < p.nphases; ++f)
    ITS_REQUIRE(p.splits <= p.phase[f].nkb, "its_conv_igemm: splits=%d exceeds the %d k-blocks of phase %d", p.splits, p.phase[f].nkb, f);
  ITS_REQUIRE(p.w_pitch % 8 == 0 && p.w_batch_stride % 8 == 0, "its_conv_igemm: weight pitch alignment");
  ITS_REQUIRE((reinterpret_cast<uintptr_t>(p.w) & 15) == 0, "its_conv_igemm: weight pointer alignment");
  ITS_REQUIRE(p.bw * p.bh * p.bb == BM, "its_conv_igemm: Hm=%d Wm=%d do not tile into 128-row boxes", p.Hm, p.Wm);
  ITS_REQUIRE(p.Wm % p.bw == 0 && (p.Hm % p.bh == 0 || p.bh > p.Hm),
              "its_conv_igemm: Hm=%d Wm=%d not divisible by the tile box", p.Hm, p.Wm);
  ITS_REQUIRE(p.w_batch_stride == 0 || p.bb == 1, "its_conv_igemm: per-image weights with a multi-image tile");
  ITS_REQUIRE(p.out_nchw || (p.out_c_pitch % 8 == 0 && (p.res == nullptr || p.res_c_pitch % 8 == 0)),
              "its_conv_igemm: out/res pitch alignment");

  // cluster size along M for the weight multicast: consecutive M tiles share the weight tile
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_b;
  // measured (scripts/conv_timeline.py): the mainloop already runs at the MMA rate with unicast
  // weights, and a cluster launch costs ~0.4 us of setup, so multicast is opt-in
  int cs = (d->cluster > 0) ? d->cluster : 1;
  if (p.w_batch_stride != 0) cs = 1;            // per-image B operands are not shared between M tiles
  ITS_REQUIRE((cs == 1 || cs == 2 || cs == 4) && tiles_m % cs == 0, "its_conv_igemm: cluster=%d does not divide %d M tiles", cs, tiles_m);
  ITS_REQUIRE(bn % (8 * cs) == 0, "its_conv_igemm: bn=%d not divisible into %d multicast slices", bn, cs);

  CUtensorMap tmA[ITS_MAX_SRC], tmB;
  {
    int rc = tapgemm_encode_operand_maps(p, bn / cs, tmA, &tmB, 1);
    if (rc != ITS_OK) return rc;
  }
  switch (bn) {
    case 32:  return launch_cs<32, 4, 2>(cs, p, tmA, tmB, stream);
    case 64:  return launch_cs<64, 4, 2>(cs, p, tmA, tmB, stream);
    case 128: return launch_cs<128, 3, 2>(cs, p, tmA, tmB, stream);
    case 192: return launch_cs<192, 5, 1>(cs, p, tmA, tmB, stream);
    default:  return launch_cs<256, 4, 1>(cs, p, tmA, tmB, stream);
  }
}

}  // namespace its
