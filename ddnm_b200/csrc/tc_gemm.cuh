// tcgen05 implicit-GEMM convolution / GEMM with 3x fp16 split ("fp32-grade" products on the
// 5th-gen tensor cores).  See tc_gemm.cu for the kernel; this header is the host-side launch record.
#pragma once
#include "common.cuh"

namespace ddnm {

enum TcTapMode : int {
  TAPS_3X3 = 0,     // 3x3, stride 1, zero pad 1 (pad comes from TMA out-of-bounds zero fill)
  TAPS_1X1 = 1,     // 1x1 / plain GEMM rows
  TAPS_3X3_S2 = 2,  // 3x3, stride 2, pad (0,1,0,1): source is stored as 4 parity phases (space-to-depth)
  TAPS_UP2X2 = 3,   // one output parity phase of (nearest x2 upsample -> 3x3 pad 1): a 2x2 stencil on the LOW-res source
};

struct TcParams {
  int H, W, N;                 // OUTPUT spatial size and number of images
  int bw, bh, bn;              // 128-row M tile = bn images x bh rows x bw columns
  int tiles_x, tiles_y, tiles_n;
  int n_tiles;                 // Cout / BN
  int mode0;                   // TcTapMode of source 0
  int cb0, kb0;                // source 0: 64-channel blocks per tap, total k-blocks (= taps * cb0)
  int kb1;                     // source 1 (always 1x1, e.g. the nin_shortcut input): k-blocks, 0 = absent
  int phase_stride;            // TAPS_3X3_S2: images per parity phase in the source's outer dim
  int up_py, up_px;            // TAPS_UP2X2: output parity (row, column) this launch produces
  int b_batched;               // 1: B has one matrix per image (3D map, z = image); 2: B is a 4D map sharing the A tile's
                               //    two outer coordinates (attention: y = head, n = image)
  long long out_sn, out_sy, out_sx;  // output element strides per image / row / column of the M tile's pixel grid
  int Cout, ldc;
  float* out;                  // out[pixel*ldc + co]
  const float* chanadd;        // chanadd[image*ca_ld + co] added per (image, channel): bias (+ timestep projection); may be null
  int ca_ld;                   // 0 = one row broadcast over images
  const float* residual;       // residual[pixel*ldr + co]; may be null
  int ldr;
  int res_mode;                // 0: same pixel; 1: nearest-upsampled source (H/2 x W/2); 2: 2x2 average of a (2H x 2W) source
  float alpha;                 // out = alpha*acc + chanadd + residual
  StatAcc* stats;              // optional GroupNorm sums of the OUTPUT: stats[(image*st_ld + co)*2 + {0,1}] += {sum, sumsq}
  int st_ld;
  int split_k = 1;             // > 1: this many CTAs per tile, each over its own k-block range, partial sums to out + ks * split_stride
  long long split_stride = 0;  //      (single-CTA forms only; no chanadd / residual / stats: splitk_reduce applies them)
  int deal = 0;                // tile -> CTA map: 0 round-robin, 1 one contiguous range per CTA (see conv_tc_kernel)
  int terms;                   // 3: hi*hi + hi*lo + lo*hi (fp32-grade, default); 1: hi*hi only (plain fp16 inputs, fast mode)
  uint32_t desc_hi;            // UMMA smem descriptor high word (SW128 K-major), see tc_gemm.cu
  uint32_t idesc;              // UMMA instruction descriptor
};

struct TcLaunch {
  CUtensorMap a0h, a0l, a1h, a1l, bh, bl, b2;   // b2: B_hi with a BN/2-row box (PAIR + DUAL form), else = bh
  TcParams p;
  int BN = 128;
  bool pair = false;           // CTA-pair kernel (cta_group::2, 256-row MMAs, cluster of 2)
  bool halo = false;           // halo-row form: A staged once per (channel slice, dy) and shared by the three dx taps
  bool dual = false;           // two partial accumulators per stage: hi*hi and hi*lo issued as one N = 2*BN instruction
  int grid = 0;
  double flops = 0;            // algorithmic flops (2*M*N*K, counted once)
};

// Build the launch record.  src0/src1: fp16 split activations; w_hi/w_lo: [batch][Cout][Ktot] fp16 K-major with
// Ktot = taps*C0 + C1 (k index = tap*C0 + ci, then source-1 channels).
TcLaunch tc_make_launch(const SplitView& src0, int mode0, const SplitView* src1, const __half* w_hi, const __half* w_lo,
                        int w_batches, int Cout, const View& out, const float* chanadd, int ca_ld, const float* residual,
                        int ldr, float alpha, int num_sms, int res_mode = 0);
void tc_run(const TcLaunch& L, cudaStream_t stream);
// One parity phase (py, px) of conv3x3(nearest_upsample_x2(src)): src is the LOW-res split, w_* the phase's pre-summed
// [Cout][4*Cin] weights (see presum_up2_weights), out the FULL-res view; writes out[:, 2y+py, 2x+px, :].
TcLaunch tc_make_up2_launch(const SplitView& src, const __half* w_hi, const __half* w_lo, int Cout, const View& out, const float* chanadd,
                            int ca_ld, int py, int px, int num_sms);

// Strided fp16 (hi, lo) operand for the batched-GEMM builder: element (k, row, head, image) at
// base[k + row*s_row + head*s_head + image*s_img]; k extent = K (multiple of 64).
struct GemmOperand {
  const __half* hi;
  const __half* lo;
  long long s_row, s_head, s_img;
};
// out[img*out_sn + head*out_sy + m*out_sx + n] = alpha * sum_k A[k, m, head, img] * B[k, n, head, img]
// (multi-head attention: QK^T and PV).  M % 128 == 0, N % 64 == 0, K % 64 == 0.
TcLaunch tc_make_gemm_launch(const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int heads, int images, float* out,
                             long long out_sn, long long out_sy, long long out_sx, float alpha, int num_sms);

// ---- fused GroupNorm + SiLU + split + 3x3 convolution (tc_gn_conv.cu): the A operand is produced inside the kernel ----
struct GnAffine {
  const float* gamma = nullptr;   // nullptr: no normalisation (raw split)
  const float* beta = nullptr;
  float eps = 1e-6f;
  int groups = 32;
  bool silu = true;
  const float* ss = nullptr;      // optional per-(image, channel) [scale(C) | shift(C)] rows (use_scale_shift_norm, unet.py:250-252)
  int ss_ld = 0;
};
struct TcGnParams {
  TcParams t;                     // tiling, B operand, epilogue (tile = 128 pixels of one row; pairs only)
  const float* x;                 // fp32 NHWC source of the 3x3 taps, row pitch x_ld
  int x_ld;
  const float* xs;                // optional fp32 NHWC source of the 1x1 side input (nin_shortcut), row pitch xs_ld
  int xs_ld;
  const StatAcc* st_in;           // per-(image, channel) sums of x (View::st)
  int st_ld_in;
  const float* gamma;
  const float* beta;
  float eps;
  int groups;
  const float* ss;
  int ss_ld;
  int silu, norm;
  int desc_mode;                  // 0: shifted start address only; 1: + base-offset field (descriptor bits [49,52))
  int pf_dist;                    // L2 prefetch distance in units (0 = none)
  long long* dbg;                 // optional (tests/diag): per-CTA clock counters, 16 per CTA — see conv_gn_tc_kernel
};
struct TcGnLaunch {
  CUtensorMap bh, bl, b2;
  TcGnParams g;
  int BN = 128;
  int grid = 0;
  double flops = 0;
};
// x: fp32 activation with its GroupNorm sums; side: raw fp32 input of a 1x1 shortcut riding as extra K blocks (may be null);
// w_hi/w_lo as for tc_make_launch with Ktot = 9*x.C + side.C.
bool tc_gn_eligible(const View& x, const View* side, int Cout, const View& out);
TcGnLaunch tc_make_gn_launch(const View& x, const GnAffine& gn, const View* side, const __half* w_hi, const __half* w_lo, int Cout,
                             const View& out, const float* chanadd, int ca_ld, const float* residual, int ldr, int num_sms);
void tc_gn_run(const TcGnLaunch& L, cudaStream_t stream);
void tc_debug_gn_desc_mode(int mode);   // tests: how the shifted A start address is described to the tensor core
void tc_debug_gn_counters(long long* dev_buf);   // diag: where the fused kernel's warps spend their clocks (nullptr = off)
void tc_debug_gn_pf_dist(int d);       // diag: L2 prefetch distance of fused launches built afterwards
void tc_debug_gn_fused(int on);         // 1: eligible layers use the fused kernel; 0 (default): gn_apply + conv_tc

// debug knobs (tests only): override descriptor words for the NEXT launches built
void tc_debug_override(uint32_t desc_hi, uint32_t idesc_xor);
void tc_debug_force_bn(int bn);
void tc_debug_halo(int on);         // 1 (default, env DDNM_HALO): eligible pair launches use the halo-row form
void tc_debug_deal(int mode);        // -1 (default): contiguous tile ranges where they pay (one N tile + GroupNorm sums), 0 / 1: force
void tc_debug_pair_dual(int on);     // 1 (default): CTA pairs at BN = 128 use the PAIR + DUAL form
void tc_debug_dual_mode(int mode);   // 1: DUAL kernel for single-CTA BN <= 128 launches (default), 0: never
void tc_debug_pair_mode(int mode);   // -1: cost model decides (default), 0: never, 1: CTA pairs wherever legal
// number of fp16 product terms used by launches built from now on (3 = parity mode, 1 = fast mode)
void tc_set_terms(int terms);
int tc_get_terms();

}  // namespace ddnm
