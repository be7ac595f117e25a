// Fused GroupNorm + SiLU + fp16 hi/lo split + 3x3 convolution on the tcgen05 tensor cores (sm_100a).
//
//   out = conv3x3( silu( groupnorm(x) ) ) [+ conv1x1(x_side)] + chanadd + residual
//
// What the unfused path does in two kernels with an HBM round trip in between (gn_apply_kernel writes the fp16 hi/lo planes,
// conv_tc_kernel reads them back 9 times through L2), this kernel does in one: the A operand of the implicit GEMM is produced
// INSIDE the convolution kernel.  Eight "transform" warps (two groups taking alternate units) read the raw fp32 activation rows
// straight from global memory into registers, apply the per-(image, channel) GroupNorm affine + SiLU (computed from the producer's running sums, View::st),
// split to fp16 hi/lo and store the 128B-swizzled K-major operand rows into shared memory — ONE halo row of 130 pixels per
// (row offset dy, 64-channel slice), which then feeds the THREE taps dx = -1, 0, +1 through UMMA descriptors whose start
// address is shifted by one 128-byte operand row per tap.  So per element the normalisation runs 3x (once per dy) instead of
// being materialised, the A operand is filled into shared memory 3x less often than with one TMA box per tap, and the planes
// never exist in HBM.
//
// Replaces, on the reference path,  h = self.conv1(nonlinearity(self.norm1(x)))  /  h = self.conv2(nonlinearity(self.norm2(h)))
// + nin_shortcut(x)  (guided_diffusion/models.py:115-134) and the in_layers / out_layers convolutions of the plain ResBlock
// (guided_diffusion/unet.py:195-252), for the layers whose rows are at least 128 pixels wide (256x256 and 128x128 maps).
//
// CTA pair (cluster of 2, tcgen05 cta_group::2, 256-row MMAs): each CTA owns one tile of 128 consecutive pixels of an image
// row.  14 warps per CTA: warp 0 = TMA producer of the weight (B) tiles, warp 1 = UMMA issuer (leader CTA) + TMEM owner,
// warps 2-5 = epilogue (TMEM -> registers -> +bias/temb/residual -> HBM, GroupNorm sums of the OUTPUT), warps 6-13 = transform.
//   BN = 128: PAIR + DUAL form — A_hi x [B_hi; B_lo] as one 256 x 256 instruction (leader's smem holds the B_hi plane, the peer's
//             the B_lo plane) + A_lo x B_hi as a 256 x 128 instruction; two partial accumulators summed by the epilogue.
//   BN = 256: plain pair form — hi*hi, hi*lo, lo*hi as three 256 x 256 instructions, each CTA staging half of the B rows.
#include <algorithm>
#include <cstdlib>

#include "tc_gemm.cuh"

namespace ddnm {

static constexpr int GK = 64;                         // fp16 elements per k-block = one 128-byte swizzle row
static constexpr int G_AROWS = 136;                    // 130 halo pixels (x0-1 .. x0+128), padded to a multiple of 8 rows
static constexpr int G_APLANE = G_AROWS * 128;         // 17 KiB, a multiple of 1024 (swizzle pattern alignment)
static constexpr int G_AUNIT = 2 * G_APLANE;           // hi plane + lo plane
static constexpr int G_NA = 3;                         // A ring depth (units)
static constexpr int G_MAXC = 512;                     // widest normalised input (coefficient table in shared memory)
static constexpr int G_THREADS = 14 * 32;

template <int BN>
struct GnCfg {
  static constexpr bool PD = BN == 128;
  // B regions of a stage.  PD: X = a FULL BN-row plane (B_hi in the leader, B_lo in the peer), Y = this CTA's BN/2-row half of
  // B_hi.  BN = 256: X = this CTA's half of B_hi, Y = its half of B_lo.
  static constexpr int BX = PD ? BN * GK * 2 : (BN / 2) * GK * 2;
  static constexpr int BY = (BN / 2) * GK * 2;
  static constexpr int B_STAGE = BX + BY;
  static constexpr int NB = PD ? 4 : 3;
  static constexpr int ACC_COLS = 256;                 // PD: two partial accumulators of 128 columns; BN = 256: one of 256
  static constexpr int TMEM_COLS = 512;
  static constexpr int RING_BYTES = G_NA * G_AUNIT + NB * B_STAGE;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * G_MAXC * 8 /*coefficient tables of the two transform groups*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory capacity");
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// arrive (release, cluster scope) on the LEADER CTA's copy of a barrier: the transform warps of both CTAs publish their operand rows
__device__ __forceinline__ void mbar_arrive_leader_release(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("ddnm_b200: mbarrier (cluster) wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// 2^x on the special-function unit (x * sigmoid(x) = x / (1 + 2^(-x log2 e)); relative error ~1e-6, far inside the parity tolerance)
__device__ __forceinline__ float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int BN>
__global__ void __launch_bounds__(G_THREADS, 1)
conv_gn_tc_kernel(const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl,
                  const __grid_constant__ CUtensorMap tm_b2, const TcGnParams g) {
  using Cfg = GnCfg<BN>;
  constexpr bool PD = Cfg::PD;
  constexpr int NB = Cfg::NB;
  const TcParams& p = g.t;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = smem_base + G_NA * G_AUNIT;
  const uint32_t bar_base = smem_base + Cfg::RING_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (G_NA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * G_NA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * G_NA + NB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * G_NA + 2 * NB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * G_NA + 2 * NB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * G_NA + 2 * NB + 4);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  float2* coef = reinterpret_cast<float2*>(gen_base + Cfg::RING_BYTES + 256);

  pdl_prologue();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cb0 = p.cb0;                       // 64-channel slices of the normalised 3x3 input
  const int upt = 3 * cb0 + p.kb1;             // A units per tile: (slice, dy) rows + 1x1 side slices
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // p.deal: one contiguous range of tile pairs per cluster (consecutive rows of one image: the output's GroupNorm sums are
  // flushed once or twice per CTA, the coefficient table is rebuilt as rarely) instead of the round-robin order
  const int n_units = total_tiles / 2, n_workers = (int)(gridDim.x / 2), worker = (int)cluster_id_x();
  const int unit_begin = p.deal ? (int)((long long)worker * n_units / n_workers) : worker;
  const int unit_end = p.deal ? (int)((long long)(worker + 1) * n_units / n_workers) : n_units;
  const int unit_step = p.deal ? 1 : n_workers;
  auto tile_of = [&](int u) {
    const int mp = u / p.n_tiles;
    return (2 * mp + (int)rank) * p.n_tiles + (u - mp * p.n_tiles);
  };
  // tile = 128 consecutive pixels of one image row (bw = 128, bh = 1, bn = 1)
  auto decode = [&](int tile, int& n_idx, int& x0, int& y, int& n) {
    n_idx = tile % p.n_tiles;
    int m = tile / p.n_tiles;
    const int tx = m % p.tiles_x;
    m /= p.tiles_x;
    y = m % p.tiles_y;
    n = m / p.tiles_y;
    x0 = tx * 128;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_bh);
    tma_prefetch_desc(&tm_bl);
    tma_prefetch_desc(&tm_b2);
    for (int s = 0; s < G_NA; ++s) {
      mbar_init(a_full(s), 8);                 // one arrival per warp of the transform GROUP that filled the unit, both CTAs (leader's copy)
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);           // every epilogue thread of both CTAs arrives once per tile
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------ weight (B) producer ------------------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t stage_tx = 2u * (uint32_t)Cfg::B_STAGE;       // both CTAs' bytes land on the leader's barrier
      for (int u = unit_begin; u < unit_end; u += unit_step) {
        int n_idx, x0, y, n;
        decode(tile_of(u), n_idx, x0, y, n);
        for (int j = 0; j < upt; ++j) {
          const bool side = j >= 3 * cb0;
          const int c = side ? j - 3 * cb0 : j / 3;
          const int dyi = side ? 0 : j - 3 * c;                     // 0..2 <-> dy = -1..1
          const int ntap = side ? 1 : 3;
          for (int dxi = 0; dxi < ntap; ++dxi) {
            const int kb = side ? p.kb0 + c : ((dyi * 3 + dxi) * cb0 + c);
            mbar_wait(b_empty(stage), phase ^ 1u);
            const uint32_t sb = b_ring + stage * Cfg::B_STAGE;
            const uint32_t fb = b_full(stage);
            if (leader) mbar_expect_tx(fb, stage_tx);
            if (PD) {
              tma_load_3d_pair(sb, rank == 0 ? &tm_bh : &tm_bl, fb, kb * GK, n_idx * BN, 0);
              tma_load_3d_pair(sb + Cfg::BX, &tm_b2, fb, kb * GK, n_idx * BN + (int)rank * (BN / 2), 0);
            } else {
              const int brow = n_idx * BN + (int)rank * (BN / 2);
              tma_load_3d_pair(sb, &tm_bh, fb, kb * GK, brow, 0);
              tma_load_3d_pair(sb + Cfg::BX, &tm_bl, fb, kb * GK, brow, 0);
            }
            if (++stage == NB) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ UMMA issuer -------------------------------------------------
    if (lane == 0 && leader) {
      uint32_t sb_i = 0, b_phase = 0, ua = 0, a_phase = 0, acc = 0, acc_phase = 0;
      long long w_a = 0, w_b = 0, w_t = 0, t_all = g.dbg ? clock64() : 0;
      const bool prof = g.dbg != nullptr;
      const uint32_t idesc_wide = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)(256 >> 3) << 17);   // N = 256
      for (int u = unit_begin; u < unit_end; u += unit_step) {
        long long q0 = prof ? clock64() : 0;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        if (prof) w_t += clock64() - q0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        bool first = true;
        for (int j = 0; j < upt; ++j) {
          const bool side = j >= 3 * cb0;
          const int ntap = side ? 1 : 3;
          q0 = prof ? clock64() : 0;
          mbar_wait_cluster(a_full(ua), a_phase);
          if (prof) w_a += clock64() - q0;
          tc_fence_after();
          const uint32_t au = a_ring + ua * G_AUNIT;
          for (int dxi = 0; dxi < ntap; ++dxi) {
            q0 = prof ? clock64() : 0;
            mbar_wait(b_full(sb_i), b_phase);
            if (prof) w_b += clock64() - q0;
            tc_fence_after();
            // A rows: main units hold pixels x0-1 .. x0+128 in rows 0..129, tap dx reads rows dx+1 .. dx+128; side units hold
            // pixels x0 .. x0+127 in rows 0..127.  One operand row = 128 bytes, so the shift is a start-address offset; the swizzle
            // phase of a shifted start goes into the descriptor's base-offset field when desc_mode asks for it.
            const uint32_t row_off = side ? 0u : (uint32_t)dxi;
            uint32_t dhi = p.desc_hi;
            if (g.desc_mode == 1) dhi |= (row_off & 7u) << 17;       // base offset, descriptor bits [49,52)
            const uint64_t ahi_w = (uint64_t)dhi << 32;
            const uint64_t bhi_w = (uint64_t)p.desc_hi << 32;
            const uint32_t ah = (((au + row_off * 128u) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t al = (((au + G_APLANE + row_off * 128u) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t sb = b_ring + sb_i * Cfg::B_STAGE;
            const uint32_t bx = ((sb & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t by = (((sb + Cfg::BX) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
            for (int k = 0; k < GK / 16; ++k) {
              const uint32_t adv = 2u * k;
              const uint32_t accum = (first && k == 0) ? 0u : 1u;
              if (PD) {
                umma_f16_pair(d_tmem, ahi_w | (ah + adv), bhi_w | (bx + adv), idesc_wide, accum);   // [hi*hi | hi*lo]
                umma_f16_pair(d_tmem, ahi_w | (al + adv), bhi_w | (by + adv), p.idesc, 1u);         // lo*hi
              } else {
                umma_f16_pair(d_tmem, ahi_w | (ah + adv), bhi_w | (bx + adv), p.idesc, accum);      // hi*hi
                umma_f16_pair(d_tmem, ahi_w | (ah + adv), bhi_w | (by + adv), p.idesc, 1u);         // hi*lo
                umma_f16_pair(d_tmem, ahi_w | (al + adv), bhi_w | (bx + adv), p.idesc, 1u);         // lo*hi
              }
            }
            first = false;
            umma_commit_pair(b_empty(sb_i));
            if (++sb_i == NB) {
              sb_i = 0;
              b_phase ^= 1u;
            }
          }
          umma_commit_pair(a_empty(ua));     // the unit's rows (in both CTAs) may be overwritten once these MMAs have read them
          if (++ua == G_NA) {
            ua = 0;
            a_phase ^= 1u;
          }
        }
        umma_commit_pair(tfull_bar(acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
      if (prof) {
        long long* d = g.dbg + (size_t)blockIdx.x * 16;
        d[10] = clock64() - t_all; d[11] = w_a; d[12] = w_b; d[13] = w_t;
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ epilogue ----------------------------------------------------
    const int ew = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = ew * 32 + lane;            // pixel of the tile
    uint32_t acc = 0, acc_phase = 0;
    float run_s[BN / 32], run_q[BN / 32], cmp_s[BN / 32], cmp_q[BN / 32];   // (value, compensation) pairs: see conv_tc_kernel
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      run_s[ch] = 0.f;
      run_q[ch] = 0.f;
      cmp_s[ch] = 0.f;
      cmp_q[ch] = 0.f;
    }
    for (int u = unit_begin; u < unit_end; u += unit_step) {
      int n_idx, x0, y, n;
      decode(tile_of(u), n_idx, x0, y, n);
      const long long pix = ((long long)n * p.H + y) * p.W + (x0 + r);
      float* orow = p.out + (long long)n * p.out_sn + (long long)y * p.out_sy + (long long)(x0 + r) * p.out_sx + n_idx * BN;
      const float* rrow = p.residual ? p.residual + pix * p.ldr + n_idx * BN : nullptr;
      const float* crow = p.chanadd ? p.chanadd + (long long)n * p.ca_ld + n_idx * BN : nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float4 cv[8], rv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cv[j] = crow ? __ldg(reinterpret_cast<const float4*>(crow + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          rv[j] = rrow ? *reinterpret_cast<const float4*>(rrow + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t v[32];
        tmem_ld32(t0 + c0, v);
        tmem_ld_wait();
        if (PD) {   // add the hi*lo partial sums kept in the stage's second half
          uint32_t v2[32];
          tmem_ld32(t0 + BN + c0, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        float ov[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ov[4 * j + 0] = p.alpha * __uint_as_float(v[4 * j + 0]) + cv[j].x + rv[j].x;
          ov[4 * j + 1] = p.alpha * __uint_as_float(v[4 * j + 1]) + cv[j].y + rv[j].y;
          ov[4 * j + 2] = p.alpha * __uint_as_float(v[4 * j + 2]) + cv[j].z + rv[j].z;
          ov[4 * j + 3] = p.alpha * __uint_as_float(v[4 * j + 3]) + cv[j].w + rv[j].w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(orow + c0 + 4 * j) = make_float4(ov[4 * j], ov[4 * j + 1], ov[4 * j + 2], ov[4 * j + 3]);
        if (p.stats) {
          // GroupNorm statistics of the output tile: transpose-reduce the 32 rows x 32 columns this warp holds so that lane L ends
          // with the column-(c0+L) sum and sum of squares over the warp's 32 pixels
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sq[j] = ov[j] * ov[j];
#pragma unroll
          for (int k = 16; k >= 1; k >>= 1) {
            const bool up = (lane & k) != 0;
#pragma unroll
            for (int i = 0; i < k; ++i) {
              const float keep_s = up ? ov[i + k] : ov[i], send_s = up ? ov[i] : ov[i + k];
              const float keep_q = up ? sq[i + k] : sq[i], send_q = up ? sq[i] : sq[i + k];
              ov[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, k);
              sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, k);
            }
          }
          two_sum_acc(run_s[c0 >> 5], cmp_s[c0 >> 5], ov[0]);
          two_sum_acc(run_q[c0 >> 5], cmp_q[c0 >> 5], sq[0]);
        }
      }
      if (p.stats) {
        // keep running sums while this CTA's consecutive tiles stay in the same image / channel block (static tile -> CTA map, so
        // the fp32 partial sums are the same every run); flush with order-independent fixed-point adds otherwise
        int next_n = -1, next_nidx = -1;
        if (u + unit_step < unit_end) {
          int nx0, ny;
          decode(tile_of(u + unit_step), next_nidx, nx0, ny, next_n);
        }
        if (next_n != n || next_nidx != n_idx) {
#pragma unroll
          for (int ch = 0; ch < BN / 32; ++ch) {
            StatAcc* d = p.stats + ((size_t)n * p.st_ld + n_idx * BN + ch * 32 + lane) * 2;
            stat_add(d, run_s[ch]);
            stat_add(d, cmp_s[ch]);
            stat_add(d + 1, run_q[ch]);
            stat_add(d + 1, cmp_q[ch]);
            run_s[ch] = 0.f;
            run_q[ch] = 0.f;
            cmp_s[ch] = 0.f;
            cmp_q[ch] = 0.f;
          }
        }
      }
      tc_fence_before();
      mbar_arrive_leader(tempty_bar(acc));
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ------------------------------------------------ transform warps ---------------------------------------------
    // Unit = one operand row block: (64-channel slice c, row offset dy) -> 130 halo pixels x 64 channels, or a 1x1 side slice ->
    // 128 pixels x 64 channels.  The 8 warps form TWO groups of 4 that take alternate units: a group issues all loads of its unit
    // (thread <-> 4 channels = one coalesced 16-byte load per row, 17 rows: row = i*8 + tw*2 + (lane >> 4)), converts, stores,
    // fences and publishes; while it waits for memory the other group converts ITS unit.  (A register double buffer inside one
    // group does not work: the generic->async proxy fence that must precede the publish is a MEMBAR that waits for the thread's
    // outstanding global loads, i.e. for the prefetch — measured with ncu: stall_membar / long-scoreboard on the fence.)
    const int grp = (warp - 6) >> 2;         // 0 / 1
    const int tw = (warp - 6) & 3;
    const int half = lane >> 4, c4 = lane & 15;
    const int gt = threadIdx.x - (6 + 4 * grp) * 32;   // 0..127 within the group
    constexpr int NIT = 17;
    const int row0 = tw * 2 + half;          // + 8*i   (so row & 7 == row0 & 7 for every row of this thread)
    float2* gcoef = coef + grp * G_MAXC;     // per-group coefficient table (the groups may be in different images)
    long long c_empty = 0, c_load = 0, c_conv = 0, c_pub = 0, c_units = 0;
    const bool prof = g.dbg != nullptr;
    const float kNegLog2e = -1.4426950408889634f;
    int cur_n = -1;
    int t_x0 = 0, t_y = 0, t_n = 0, t_u = -1;
    // this group's units: sequence index q = grp, grp + 2, ... over (tile, j); A ring slot q % G_NA
    int u = unit_begin, j = grp;
    while (j >= upt) { j -= upt; u += unit_step; }
    for (long long q = grp; u < unit_end; q += 2) {
      if (u != t_u) {                        // decode once per tile (integer divisions are slow)
        int n_idx;
        decode(tile_of(u), n_idx, t_x0, t_y, t_n);
        t_u = u;
      }
      const bool side = j >= 3 * cb0;
      const int c = side ? j - 3 * cb0 : j / 3;
      const int yy = side ? t_y : t_y + (j - 3 * c) - 1;
      const int nrows = side ? 128 : 130;
      const int px0 = side ? t_x0 : t_x0 - 1;
      const int ld = side ? g.xs_ld : g.x_ld;
      const float* base = (yy >= 0 && yy < p.H) ? (side ? g.xs : g.x) + ((long long)t_n * p.H + yy) * p.W * ld + c * GK + c4 * 4 : nullptr;
      // ---- all loads of the unit in flight
      float4 buf[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int r = i * 8 + row0;
        const int px = px0 + r;
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (base != nullptr && r < nrows && px >= 0 && px < p.W) buf[i] = ldg_nc_f4(base + (long long)px * ld);
      }
      // ---- per-(image, channel) affine of the GroupNorm (+ scale-shift) from the producer's running sums: y = a*x + b
      if (!side && t_n != cur_n) {
        named_bar_sync(1 + grp, 128);        // the group is done with the previous image's table
        const int C = cb0 * GK;
        for (int ch = gt; ch < C; ch += 128) {
          float a = 1.f, b = 0.f;
          if (g.norm) {
            const int cpg = C / g.groups;
            const int g0 = (ch / cpg) * cpg;
            double s1 = 0, s2 = 0;
            for (int k = 0; k < cpg; ++k) {
              const StatAcc* sp = g.st_in + ((size_t)t_n * g.st_ld_in + g0 + k) * 2;
              s1 += stat_value(sp[0]);
              s2 += stat_value(sp[1]);
            }
            const double cnt = (double)p.H * p.W * cpg;
            const double mean = s1 / cnt;
            double var = s2 / cnt - mean * mean;
            var = var < 0 ? 0 : var;
            const float rstd = (float)(1.0 / sqrt(var + (double)g.eps));
            a = rstd * g.gamma[ch];
            b = g.beta[ch] - (float)mean * a;
            if (g.ss) {                      // h = norm(h) * (1 + scale) + shift   (unet.py:250-252)
              const float one_plus = 1.0f + g.ss[(size_t)t_n * g.ss_ld + ch];
              a *= one_plus;
              b = fmaf(b, one_plus, g.ss[(size_t)t_n * g.ss_ld + C + ch]);
            }
          }
          gcoef[ch] = make_float2(a, b);
        }
        named_bar_sync(1 + grp, 128);
        cur_n = t_n;
      }
      float a4[4] = {1.f, 1.f, 1.f, 1.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!side) {
        const float4 ab0 = *reinterpret_cast<const float4*>(&gcoef[c * GK + c4 * 4]);
        const float4 ab1 = *reinterpret_cast<const float4*>(&gcoef[c * GK + c4 * 4 + 2]);
        a4[0] = ab0.x; b4[0] = ab0.y; a4[1] = ab0.z; b4[1] = ab0.w;
        a4[2] = ab1.x; b4[2] = ab1.y; a4[3] = ab1.z; b4[3] = ab1.w;
      }
      const bool act = !side && g.silu;
      const uint32_t ua = (uint32_t)(q % G_NA), a_phase = (uint32_t)((q / G_NA) & 1);
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (prof) t0 = clock64();
      mbar_wait(a_empty(ua), a_phase ^ 1u);
      if (prof) {
        t1 = clock64();
        uint32_t sink;   // touch the last-issued load: the scoreboard wait for the unit's loads lands here, not in the conversion
        asm volatile("mov.b32 %0, %1;" : "=r"(sink) : "f"(buf[NIT - 1].x + buf[0].x));
        t2 = clock64() + (sink & 0u);
      }
      // 16-byte chunk (c4 >> 1) ^ (row & 7), 8-byte half (c4 & 1)
      const uint32_t hi_base = a_ring + ua * G_AUNIT + (uint32_t)row0 * 128u + (uint32_t)((((c4 >> 1) ^ (row0 & 7)) << 4) + (c4 & 1) * 8);
      const uint32_t lo_base = hi_base + G_APLANE;
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int r = i * 8 + row0;
        if (r < nrows) {
          // padding pixels / rows: the ACTIVATED value is zero (conv zero padding), whatever silu(b) would be
          const int px = px0 + r;
          if (base == nullptr || px < 0 || px >= p.W) {
            st_shared_v2(hi_base + (uint32_t)i * 1024u, 0u, 0u);
            st_shared_v2(lo_base + (uint32_t)i * 1024u, 0u, 0u);
            continue;
          }
          float t[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            t[k] = fmaf(t[k], a4[k], b4[k]);
            if (act) {
              const float e = exp2f_fast(t[k] * kNegLog2e);
              t[k] = __fdividef(t[k], 1.0f + e);
            }
          }
          uint32_t h01, h23, l01, l23;
          split2_f16(t[0], t[1], h01, l01);
          split2_f16(t[2], t[3], h23, l23);
          st_shared_v2(hi_base + (uint32_t)i * 1024u, h01, h23);
          st_shared_v2(lo_base + (uint32_t)i * 1024u, l01, l23);
        }
      }
      if (prof) t3 = clock64();
      fence_proxy_async_smem();              // generic-proxy stores -> visible to the tensor core's (async proxy) reads
      __syncwarp();
      // plain (release.cta) arrive: a .release.cluster arrive adds MEMBAR.ALL.CTA + ERRBAR
      if (lane == 0) mbar_arrive_leader(a_full(ua));
      if (prof) {
        c_empty += t1 - t0; c_load += t2 - t1; c_conv += t3 - t2; c_pub += clock64() - t3; ++c_units;
      }
      j += 2;
      while (j >= upt) { j -= upt; u += unit_step; }
    }
    if (prof && tw == 0 && lane == 0) {
      long long* d = g.dbg + (size_t)blockIdx.x * 16 + grp * 5;
      d[0] = c_units; d[1] = c_empty; d[2] = c_load; d[3] = c_conv; d[4] = c_pub;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// -------------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}
static long long* g_gn_dbg = nullptr;
void tc_debug_gn_counters(long long* dev_buf) { g_gn_dbg = dev_buf; }
static int g_gn_desc_mode = env_int("DDNM_GN_DESC_MODE", 0);
static int g_gn_pf_dist = env_int("DDNM_GN_PF_DIST", 0);
void tc_debug_gn_pf_dist(int d) { g_gn_pf_dist = d; }
void tc_debug_gn_desc_mode(int mode) { g_gn_desc_mode = mode; }
// Default OFF (profiles/r02_forward_speedup.md): per launch it beats gn_apply + conv_tc on wide inputs without a side operand
// (up.0.*.conv1: 1811 us vs 1353 + 583), loses wherever the 1x1 shortcut's raw input has to be split by the transform warps as well
// (conv2+nin: 1512 vs 854 + 212) — every wide block of the celeba network; whole forward 27.8-28.0 vs 26.6-27.3 ms (celeba, B = 16),
// 46.8 vs 47.7 ms (imagenet, B = 8).  The halo-row form of conv_tc_kernel has since taken over its A-operand reuse.
static int g_gn_enable = env_int("DDNM_GN_FUSED", 0);
void tc_debug_gn_fused(int on) { g_gn_enable = on; }
bool tc_gn_enabled() { return g_gn_enable != 0; }

static bool gn_shape_ok(const View& x, const View* side, int Cout, const View& out) {
  if (out.W % 128 != 0 || x.C % GK != 0 || x.C > G_MAXC || Cout % 128 != 0) return false;
  if (x.H != out.H || x.W != out.W || x.N != out.N) return false;
  if (((long long)out.N * out.H * (out.W / 128)) % 2 != 0) return false;
  if (x.ld % 4 != 0 || ((uintptr_t)x.p & 15) != 0) return false;
  if (side && (side->C % GK != 0 || side->ld % 4 != 0 || ((uintptr_t)side->p & 15) != 0 || side->H != out.H || side->W != out.W)) return false;
  return true;
}

bool tc_gn_eligible(const View& x, const View* side, int Cout, const View& out) { return g_gn_enable != 0 && gn_shape_ok(x, side, Cout, out); }

CUtensorMap tc_make_weight_map(const __half* w, int Ktot, int Cout, int box_rows);   // tc_gemm.cu

TcGnLaunch tc_make_gn_launch(const View& x, const GnAffine& gn, const View* side, const __half* w_hi, const __half* w_lo, int Cout,
                             const View& out, const float* chanadd, int ca_ld, const float* residual, int ldr, int num_sms) {
  DDNM_CHECK(gn_shape_ok(x, side, Cout, out), "fused GroupNorm convolution: unsupported shape");
  TcGnLaunch L;
  TcParams& p = L.g.t;
  p.H = out.H; p.W = out.W; p.N = out.N;
  p.bw = 128; p.bh = 1; p.bn = 1;
  p.tiles_x = out.W / 128; p.tiles_y = out.H; p.tiles_n = out.N;
  L.BN = (Cout % 256 == 0) ? 256 : 128;
  p.n_tiles = Cout / L.BN;
  p.mode0 = TAPS_3X3;
  p.cb0 = x.C / GK;
  p.kb0 = 9 * p.cb0;
  p.kb1 = side ? side->C / GK : 0;
  p.phase_stride = 0; p.up_py = p.up_px = 0; p.b_batched = 0;
  p.Cout = Cout; p.ldc = out.ld; p.out = out.p;
  p.out_sx = out.ld; p.out_sy = (long long)out.W * out.ld; p.out_sn = (long long)out.H * out.W * out.ld;
  DDNM_CHECK(out.C == Cout && out.ld % 4 == 0 && ((uintptr_t)out.p & 15) == 0, "output view misaligned");
  p.chanadd = chanadd; p.ca_ld = ca_ld; p.residual = residual; p.ldr = ldr; p.alpha = 1.0f; p.res_mode = 0;
  if (residual) DDNM_CHECK(ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0, "residual misaligned");
  p.stats = out.st; p.st_ld = out.st_ld;
  p.terms = 3;
  DDNM_CHECK(tc_get_terms() == 3, "the fused GroupNorm convolution implements the fp32-grade (3-term) arithmetic only");
  p.desc_hi = 64u | (1u << 14) | (2u << 29);
  p.idesc = (1u << 4) | ((uint32_t)(L.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  TcGnParams& g = L.g;
  g.x = x.p; g.x_ld = x.ld;
  g.xs = side ? side->p : nullptr; g.xs_ld = side ? side->ld : 0;
  g.norm = gn.gamma != nullptr ? 1 : 0;
  if (g.norm) DDNM_CHECK(x.st != nullptr && x.C % gn.groups == 0, "normalisation needs the tensor's per-channel sums (View::st)");
  g.st_in = x.st; g.st_ld_in = x.st_ld;
  g.gamma = gn.gamma; g.beta = gn.beta; g.eps = gn.eps; g.groups = gn.groups; g.ss = gn.ss; g.ss_ld = gn.ss_ld; g.silu = gn.silu ? 1 : 0;
  g.desc_mode = g_gn_desc_mode;
  g.dbg = g_gn_dbg;
  g.pf_dist = g_gn_pf_dist;
  const int Ktot = (p.kb0 + p.kb1) * GK;
  const bool pd = L.BN == 128;
  L.bh = tc_make_weight_map(w_hi, Ktot, Cout, pd ? L.BN : L.BN / 2);
  L.bl = tc_make_weight_map(w_lo, Ktot, Cout, pd ? L.BN : L.BN / 2);
  L.b2 = pd ? tc_make_weight_map(w_hi, Ktot, Cout, L.BN / 2) : L.bh;
  const int total = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;
  L.grid = 2 * std::min(total / 2, num_sms / 2);
  p.deal = (p.n_tiles == 1 && total / 2 >= L.grid) ? 1 : 0;
  L.flops = 2.0 * (double)out.pixels() * Cout * Ktot;
  return L;
}

template <int BN>
static void launch_gn(const TcGnLaunch& L, cudaStream_t stream) {
  using Cfg = GnCfg<BN>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set))
    CUDA_CHECK(cudaFuncSetAttribute(conv_gn_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  launch_pdl(conv_gn_tc_kernel<BN>, dim3(L.grid), dim3(G_THREADS), (size_t)Cfg::SMEM_BYTES, stream, 2, L.bh, L.bl, L.b2, L.g);
  CUDA_CHECK(cudaGetLastError());
}

void tc_gn_run(const TcGnLaunch& L, cudaStream_t stream) {
  if (L.BN == 256) launch_gn<256>(L, stream);
  else launch_gn<128>(L, stream);
}

}  // namespace ddnm
