// Shared helpers for the ddnm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ddnm {

// ---------------------------------------------------------------------------------------------
// Error handling: every C-ABI entry point catches ddnm::Error and stores the message thread-locally.
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define DDNM_CHECK(cond, msg)                                                                     \
  do {                                                                                            \
    if (!(cond)) throw ::ddnm::Error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)); \
  } while (0)

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw ::ddnm::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr " at " __FILE__ ":" + \
                          std::to_string(__LINE__));                                              \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
// One-time per-DEVICE actions (cudaFuncSetAttribute is a per-device setting): true the first time `flags` sees the current device.
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}
inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// Stream-ordered scratch that is returned to the pool on every exit path (a DDNM_CHECK throw inside a sampling loop must not
// leak the loop's buffers).
struct StreamBuf {
  float* p = nullptr;
  cudaStream_t st = nullptr;
  StreamBuf(size_t elems, cudaStream_t s) : st(s) {
    cudaError_t e = cudaMallocAsync((void**)&p, elems * sizeof(float), s);
    if (e != cudaSuccess) throw Error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in cudaMallocAsync (sampler scratch)");
  }
  ~StreamBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  StreamBuf(const StreamBuf&) = delete;
  StreamBuf& operator=(const StreamBuf&) = delete;
};

// One GroupNorm running sum: a two's-complement fixed-point accumulator with 48 fractional bits kept as two carry-free words
// (value = (hi * 2^32 + lo) * 2^-48, see stat_add).  Partial sums from many CTAs are added with integer atomics, so the total
// does not depend on the order in which they arrive: a forward pass is bit-reproducible run to run (floating-point atomics
// are not).  Range +-2^47, resolution 2^-48 (what a float partial sum loses below that is far under the fp32 noise of the
// statistics themselves).
struct StatAcc {
  unsigned long long lo;
  long long hi;
};

// Programmatic dependent launch (PDL): consecutive kernels of the forward carry cudaLaunchAttributeProgrammaticStreamSerialization,
// so a kernel's CTAs may be scheduled (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) while the
// previous kernel's last CTAs drain; pdl_prologue() at the top of every such kernel (1) releases ITS successor and (2) blocks
// until the predecessor grid has completed and flushed, before any dependent global access.  Captured into the CUDA graph these
// become programmatic edges.  DDNM_PDL=0 launches everything fully serialised (A/B measurements).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess) throw Error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in cudaLaunchKernelEx");
}

// A strided NHWC fp32 activation view: element (n, y, x, c) at p[((n*H + y)*W + x)*ld + c].
// ld >= C lets a tensor live inside a channel slice of a wider (concat) buffer.
// st (optional): per-(image, channel) running sums for GroupNorm, st[(n*st_ld + c)*2 + {0,1}] = {sum, sum of squares}
// over the image's pixels; filled by the kernel that PRODUCES the tensor (tensor-core epilogue / gn_stats), so the
// normalisation never re-reads the tensor just to reduce it.
struct View {
  float* p = nullptr;
  int N = 0, H = 0, W = 0, C = 0, ld = 0;
  StatAcc* st = nullptr;
  int st_ld = 0;
  long long pixels() const { return (long long)N * H * W; }
  View slice(int c0, int c) const {
    View v = *this;
    v.p = p + c0;
    v.C = c;
    if (st) v.st = st + 2 * (size_t)c0;
    return v;
  }
};

// fp16 (hi, lo) split of an activation: two contiguous NHWC planes.
struct SplitView {
  __half* hi = nullptr;
  __half* lo = nullptr;
  int N = 0, H = 0, W = 0, C = 0;   // dims of the planes as the consumer sees them
};

#ifdef __CUDACC__
// see launch_pdl: release the successor grid, then wait for the predecessor grid (no-ops without a programmatic dependency)
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05.  Addresses are 32-bit shared-window addresses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (-> CUDA error) after ~2 s instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("ddnm_b200: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// --- tcgen05 ---
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// 32-byte global accesses (sm_100: LDG/STG .256): one whole sector per thread
__device__ __forceinline__ void ldg_f32x8(const float* p, float* r) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_f32x8(float* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]),
               "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// --- CTA pair (cta_group::2): two SMs of one TPC execute one 256-row MMA; CTA rank 0 ("leader") issues it.  The shared::cluster
// address of a barrier with bit 24 cleared names the copy in the even (leader) CTA of the pair. ---
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads issued by either CTA of the pair; the bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the leader CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (256 rows: 128 from each CTA's smem) * B (N rows: N/2 from each CTA's smem)
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the barrier at the same shared-memory offset in BOTH CTAs of the pair gets one arrival when the issued MMAs complete
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// acc += p, independent of the order of concurrent adds.  p is converted (exactly, unless it has bits below 2^-48) to a signed
// fixed-point integer X with 48 fractional bits, and X = W1 * 2^32 + W0 (W0 = X mod 2^32 >= 0, W1 = floor(X / 2^32)) is added
// as TWO independent integer reductions: acc->lo += W0 (up to 2^32 partial sums cannot wrap it), acc->hi += W1 (range +-2^47).
// No carry crosses between the words, so neither atomic needs its return value: both compile to fire-and-forget RED
// instructions (a carry-propagating 128-bit add needs the old low word back, a round trip to L2 per statistic that the
// epilogue warps of the tensor-core kernels had to wait for).
__device__ __forceinline__ void stat_add(StatAcc* acc, float p) {
  const int bits = __float_as_int(p);
  int ex = (bits >> 23) & 0xff;
  unsigned long long m = (unsigned long long)(bits & 0x7fffff);
  if (ex) m |= 0x800000ull; else ex = 1;          // |p| = m * 2^(ex - 150)
  const int sh = ex - 150 + 48;                   // X = m * 2^sh
  unsigned long long lo = 0, hi = 0;              // |X| as a 128-bit integer
  if (sh >= 64) hi = m << min(sh - 64, 7);         // |p| >= 2^47 cannot occur for activation sums; clamp the shift
  else if (sh > 0) { lo = m << sh; hi = sh > 40 ? m >> (64 - sh) : 0ull; }
  else if (sh > -24) lo = m >> (-sh);
  if (bits < 0) {                                 // two's-complement negate
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  const unsigned long long w0 = lo & 0xffffffffull;
  const unsigned long long w1 = (hi << 32) | (lo >> 32);
  if (w0) atomicAdd(&acc->lo, w0);
  if (w1) atomicAdd(reinterpret_cast<unsigned long long*>(&acc->hi), w1);
}
// sum += v with the rounding error of the addition collected in comp (Knuth two-sum): sum + comp is the running total to ~2^-48
__device__ __forceinline__ void two_sum_acc(float& sum, float& comp, float v) {
  const float s = __fadd_rn(sum, v);
  const float bb = __fadd_rn(s, -sum);
  comp = __fadd_rn(comp, __fadd_rn(__fadd_rn(sum, -__fadd_rn(s, -bb)), __fadd_rn(v, -bb)));
  sum = s;
}
__device__ __forceinline__ double stat_value(const StatAcc& a) {
  return ((double)a.hi * 4294967296.0 + (double)a.lo) * (1.0 / 281474976710656.0);
}

// fp32 -> (hi, lo) fp16 pair: hi = rn(x) saturated to the finite fp16 range, lo = rn(x - hi).
// hi + lo carries ~22 significant bits, so hi*hi + hi*lo + lo*hi reproduces an fp32 product to ~2^-21.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  hi = __ushort_as_half(h);
  float r = x - __half2float(hi);
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(r));
  lo = __ushort_as_half(h);
}
// two fp32 values -> packed fp16 hi pair + packed fp16 lo pair (hi = rn(x) saturated to the finite range, lo = rn(x - hi))
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float2 f = __half22float2(h);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - f.y), "f"(x0 - f.x));
}

#endif  // __CUDACC__

}  // namespace ddnm
