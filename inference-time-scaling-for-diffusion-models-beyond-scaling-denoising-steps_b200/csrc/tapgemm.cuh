// Device-side parameter block shared by the tcgen05 tap-GEMM kernel
// (conv_igemm_sm100.cu) and its CUDA-core reference twin (conv_direct.cu).
#pragma once
#include "its_common.cuh"
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched at run time)

namespace its {

struct DevSrc {
  const __nv_bfloat16* ptr;  // already offset by c_off
  int c_pitch, C, H, W, stride, bcast, fp16;
};

struct DevPhase {
  int ntaps, w_k0, py, px;
  int nkb;                   // k-blocks of 64 in this phase
  int fp16_first;            // operand format of k-blocks [0, kb_switch): 1 = IEEE fp16, 0 = bf16
  int kb_switch;             // k-blocks [kb_switch, nkb) use the other format (nkb if none)
  int8_t src[ITS_MAX_TAPS], dy[ITS_MAX_TAPS], dx[ITS_MAX_TAPS];
};

struct TapGemmParams {
  DevSrc src[ITS_MAX_SRC];
  DevPhase phase[ITS_MAX_PHASES];
  int nsrc, nphases;
  int B, Hm, Wm;
  const __nv_bfloat16* w;
  int w_pitch;
  long long w_batch_stride;
  int Cout;
  void* out;                 // already offset by out_c_off
  int out_fp32, Hout, Wout, out_scale, out_c_pitch;
  const float* bias;
  const float* vec;          // already offset by vec_off
  int vec_stride;
  const float* vec2;         // already offset by vec2_off
  int vec2_stride;
  const __nv_bfloat16* res;  // already offset by res_c_off
  int res_c_pitch;
  float alpha;
  int out_nchw;
  int splits;                // split-K factor (>= 1)
  float* ws;                 // [nphases][splits][B*Hm*Wm][Cout] fp32 partial sums (tile-per-CTA schedule);
                             // persistent schedule: [flags][tile][split][128][BN]
  int ws_flag_words;         // persistent split-K: ints reserved for the flags at the head of ws
  long long* dbg;            // optional per-CTA clock stamps (diagnostics)
  float* stats;              // GroupNorm partial sums [B][stats_parts][Cout/4][2] or null (persistent kernel)
  int stats_parts;
  int out_fp16;              // 16-bit output format: 0 = bf16, 1 = IEEE fp16
  int res_fp16;              // format of the residual tensor read by the non-folded epilogues
  // GroupNorm(+Swish) of this launch's own output fused into the persistent kernel's epilogue
  void* gn_out;              // IEEE fp16 NHWC, or null
  int gn_c_pitch;
  const float* gn_gamma;
  const float* gn_beta;
  int gn_groups;
  float gn_eps;
  int gn_silu, gn_only;
  int* gn_sync;              // [gn_sync_words] monotonic arrival counters of the peer tiles of an image
  int gn_sync_words;
  int gn_peers;              // tiles an image spans (same N tile); 1 = a tile holds whole images
  // M tiling: a 128-row tile is a (bb images) x (bh rows) x (bw cols) box
  int bw, bh, bb, tiles_x, tiles_y, tiles_b;
};

// Validates the host descriptor and fills the device parameter block.
int tapgemm_build_params(const its_conv_desc* d, TapGemmParams* p, bool need_k64);
int tapgemm_launch_ref(const TapGemmParams& p, cudaStream_t stream);
int tapgemm_launch_sm100(const its_conv_desc* d, const TapGemmParams& p, cudaStream_t stream);
// persistent variant (conv_persist_sm100.cu): bf16 NHWC output by TMA store, shared weights, no split-K
bool tapgemm_persist_eligible(const its_conv_desc* d, const TapGemmParams& p);
int tapgemm_launch_persist(const its_conv_desc* d, const TapGemmParams& p, cudaStream_t stream);
// number of GroupNorm partial-sum slots per image the persistent kernel writes for this tiling
int tapgemm_stats_parts(const its_conv_desc* d, const TapGemmParams& p);
// peer-tile counters the fused GroupNorm epilogue needs for this tiling (-1 = not fusable, see its_b200.h)
int tapgemm_gn_sync_words(const its_conv_desc* d, const TapGemmParams& p);
int encode_bf16_map(CUtensorMap* tm, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr,
                    const char* what);
int tapgemm_encode_operand_maps(const TapGemmParams& p, int b_box_rows, CUtensorMap* tmA, CUtensorMap* tmB_out,
                                int row_boxes);
int device_sm_count();

}  // namespace its
