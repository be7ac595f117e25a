// Degradation operators (functions/svd_operators.py) and the fused DDNM / DDNM+ update (functions/svd_ddnm.py:57-65,
// :114-131) as image-space CUDA kernels on NCHW fp32 images.  All kernels here are HBM-bound: algorithmic traffic of
// the fused step is xt + et + noise in, x0_t + xt_next out = 5 * 4 * C*H*W bytes per image (+ y).
#include "operators.cuh"

#include <algorithm>
#include <cmath>

#include "../../include/ddnm_b200.h"
#include "api_util.cuh"
#include "kernels.cuh"

namespace ddnm {

// ------------------------------------------------------------------------------------------------------------------
// Coefficient rules shared by every Lambda / Lambda_noise in the reference (e.g. svd_operators.py:568-604), evaluated
// per singular value with the same fp32 operation order (0-dim tensors and python floats both act as fp32 scalars).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lam_coeff(float s, const PlusScalars& ps) {
  const float inv = (s == 0.f) ? 0.f : __fdiv_rn(1.0f, s);
  float lam = 1.0f;
  if (ps.active) {
    const float thr = __fmul_rn(__fmul_rn(ps.a, ps.sigma_y), inv);
    if (ps.sigma_t < thr) lam = __fdiv_rn(__fdiv_rn(__fmul_rn(__fmul_rn(s, ps.sigma_t), ps.c), ps.a), ps.sigma_y);
  }
  return lam;
}
__device__ __forceinline__ void noise_coeff(float s, const PlusScalars& ps, float& d1, float& d2) {
  const float inv = (s == 0.f) ? 0.f : __fdiv_rn(1.0f, s);
  d1 = __fmul_rn(ps.sigma_t, ps.eta);
  d2 = __fmul_rn(ps.sigma_t, ps.c);
  if (ps.active) {
    const float thr = __fmul_rn(__fmul_rn(ps.a, ps.sigma_y), inv);
    float ci = (ps.sigma_t < thr) ? 1.f : 0.f;
    d1 = __fadd_rn(__fmul_rn(d1, 1.f - ci), __fmul_rn(__fmul_rn(ci, ps.sigma_t), ps.eta));
    d2 = __fmul_rn(d2, 1.f - ci);
    ci = (ps.sigma_t > thr) ? 1.f : 0.f;
    const float t1 = __fmul_rn(ps.sigma_t, ps.sigma_t);
    const float t3 = __fmul_rn(__fmul_rn(__fmul_rn(ps.a, ps.a), ps.sy2), __fmul_rn(inv, inv));
    d1 = __fadd_rn(__fmul_rn(d1, 1.f - ci), sqrtf(__fmul_rn(ci, __fsub_rn(t1, t3))));
    d2 = __fmul_rn(d2, 1.f - ci);
    ci = (s == 0.f) ? 1.f : 0.f;
    d1 = __fadd_rn(__fmul_rn(d1, 1.f - ci), __fmul_rn(__fmul_rn(ci, ps.sigma_t), ps.eta));
    d2 = __fadd_rn(__fmul_rn(d2, 1.f - ci), __fmul_rn(__fmul_rn(ci, ps.sigma_t), ps.c));
  }
}

PlusScalars Operator::make_plus(float a, float sigma_y, float sigma_t, float eta) {
  PlusScalars p;
  p.a = a; p.sigma_y = sigma_y; p.sigma_t = sigma_t; p.eta = eta;
  p.c = (float)std::sqrt(1.0 - (double)eta * (double)eta);
  p.sy2 = (float)((double)sigma_y * (double)sigma_y);
  p.active = (a != 0.f && sigma_y != 0.f) ? 1 : 0;
  return p;
}

// x0_t = (xt - et * sqrt(1-at)) / sqrt(at)      (svd_ddnm.py:57), unfused multiply / subtract / divide
__device__ __forceinline__ float x0_from(float xt, float et, const StepScalars& sc) {
  return __fdiv_rn(__fsub_rn(xt, __fmul_rn(et, sc.sqrt_1m_at)), sc.sqrt_at);
}
// xt_next = at_next.sqrt() * x0_hat + c1 * z + c2 * et      (svd_ddnm.py:65), evaluated left to right
__device__ __forceinline__ float renoise(float x0h, float z, float et, const StepScalars& sc) {
  return __fadd_rn(__fadd_rn(__fmul_rn(sc.sqrt_atn, x0h), __fmul_rn(sc.c1, z)), __fmul_rn(sc.c2, et));
}

// ------------------------------------------------------------------------------------------------------------------
// "Local group" operators: SuperResolution (group = r x r patch of one channel, svd_operators.py:479-623) and
// Colorization (group = the 3 channels of one pixel, :627-736).  A has rank 1 per group: A g = u00 * s0 * <V[:,0], g>.
// One thread owns one group; K x K basis V sits in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int K, int MODE>  // MODE 0: SR with R = sqrt(K); MODE 1: colour (K = 3)
struct Group {
  int b;
  long long base;  // offset of element 0 inside image b
  int D, HW;
  long long yidx;
  __device__ __forceinline__ Group(long long g, int C, int Dd, long long img_elems) {
    D = Dd;
    HW = Dd * Dd;
    if (MODE == 0) {
      constexpr int R = K == 4 ? 2 : (K == 16 ? 4 : 8);
      const int yd = Dd / R;
      const int px = (int)(g % yd);
      const int py = (int)((g / yd) % yd);
      const int c = (int)((g / ((long long)yd * yd)) % C);
      b = (int)(g / ((long long)yd * yd * C));
      base = ((long long)c * Dd + (long long)py * R) * Dd + (long long)px * R;
      yidx = ((long long)b * C + c) * yd * yd + (long long)py * yd + px;
    } else {
      const int p = (int)(g % HW);
      b = (int)(g / HW);
      base = p;
      yidx = (long long)b * HW + p;
    }
  }
  __device__ __forceinline__ long long off(int k) const {
    if (MODE == 0) {
      constexpr int R = K == 4 ? 2 : (K == 16 ? 4 : 8);
      return base + (long long)(k / R) * D + (k % R);
    }
    return base + (long long)k * HW;
  }
  __device__ __forceinline__ void load(const float* p, long long img_stride, float (&v)[K]) const {
    const float* q = p + (long long)b * img_stride;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __ldg(q + off(k));
  }
  __device__ __forceinline__ void store(float* p, long long img_stride, const float (&v)[K]) const {
    float* q = p + (long long)b * img_stride;
#pragma unroll
    for (int k = 0; k < K; ++k) q[off(k)] = v[k];
  }
};

enum LocalFn : int { LF_A = 0, LF_PINV = 1, LF_PROJECT = 2, LF_LAMBDA = 3, LF_NOISE = 4, LF_STEP = 5 };

template <int K>
__device__ __forceinline__ float local_resid(const float (&x0)[K], const float* V, float u00, float s0, float yv, float (&resid)[K]) {
  float cval = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) cval = fmaf(V[k * K], x0[k], cval);  // (V^T g)[0]
  const float av = __fmul_rn(u00, __fmul_rn(s0, cval));             // U (S V^T g)
  const float r = __fsub_rn(av, yv);
  const float cc = __fmul_rn(__fmul_rn(u00, r), __fdiv_rn(1.0f, s0));  // S^+ U^T r
#pragma unroll
  for (int k = 0; k < K; ++k) resid[k] = __fmul_rn(V[k * K], cc);   // V (cc, 0, ..)
  return av;
}

template <int K, int MODE, int FN>
__global__ void __launch_bounds__(128) local_kernel(const float* __restrict__ in0, const float* __restrict__ in1, long long in1_stride,
                                                    const float* __restrict__ in2, const float* __restrict__ y,
                                                    const float* __restrict__ Vg, float u00, float s0, StepScalars sc,
                                                    float* __restrict__ out0, float* __restrict__ out1, long long groups, int C,
                                                    int D) {
  __shared__ float V[K * K];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) V[i] = Vg[i];
  __syncthreads();
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const long long img = (long long)C * D * D;
  Group<K, MODE> G(g, C, D, img);
  const PlusScalars& ps = sc.plus;
  if (FN == LF_A) {
    float x[K];
    G.load(in0, img, x);
    float cval = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) cval = fmaf(V[k * K], x[k], cval);
    out0[G.yidx] = __fmul_rn(u00, __fmul_rn(s0, cval));
  } else if (FN == LF_PINV) {
    const float cc = __fmul_rn(__fmul_rn(u00, y[G.yidx]), __fdiv_rn(1.0f, s0));
    float o[K];
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = __fmul_rn(V[k * K], cc);
    G.store(out0, img, o);
  } else if (FN == LF_PROJECT) {
    float x[K], r[K];
    G.load(in0, img, x);
    local_resid<K>(x, V, u00, s0, y[G.yidx], r);
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = __fsub_rn(x[k], r[k]);
    G.store(out0, img, x);
  } else if (FN == LF_LAMBDA) {
    float x[K], o[K];
    G.load(in0, img, x);
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float sp = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) sp = fmaf(V[k * K + j], x[k], sp);
      sp *= lam_coeff(j == 0 ? s0 : 0.f, ps);
#pragma unroll
      for (int k = 0; k < K; ++k) o[k] = fmaf(V[k * K + j], sp, o[k]);
    }
    G.store(out0, img, o);
  } else if (FN == LF_NOISE) {
    float v[K], e[K], ov[K], oe[K];
    G.load(in0, img, v);
    G.load(in1, in1_stride, e);
#pragma unroll
    for (int k = 0; k < K; ++k) { ov[k] = 0.f; oe[k] = 0.f; }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float d1, d2;
      noise_coeff(j == 0 ? s0 : 0.f, ps, d1, d2);
      const float a = v[j] * d1, b = e[j] * d2;  // raw pixels used as spectral coordinates (svd_operators.py:581-621)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        ov[k] = fmaf(V[k * K + j], a, ov[k]);
        oe[k] = fmaf(V[k * K + j], b, oe[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) ov[k] = __fadd_rn(ov[k], oe[k]);
    G.store(out0, img, ov);
  } else {  // LF_STEP: in0 = xt, in1 = et, in2 = noise, out0 = x0_t, out1 = xt_next
    float xt[K], et[K], z[K], x0[K], r[K];
    G.load(in0, img, xt);
    G.load(in1, in1_stride, et);
    G.load(in2, img, z);
#pragma unroll
    for (int k = 0; k < K; ++k) x0[k] = x0_from(xt[k], et[k], sc);
    G.store(out0, img, x0);
    local_resid<K>(x0, V, u00, s0, y[G.yidx], r);
    float xn[K];
    if (!sc.use_plus) {
#pragma unroll
      for (int k = 0; k < K; ++k) xn[k] = renoise(__fsub_rn(x0[k], r[k]), z[k], et[k], sc);
    } else {
      float L[K], nv[K], ne[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { L[k] = 0.f; nv[k] = 0.f; ne[k] = 0.f; }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float sj = j == 0 ? s0 : 0.f;
        float sp = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) sp = fmaf(V[k * K + j], r[k], sp);
        sp *= lam_coeff(sj, ps);
        float d1, d2;
        noise_coeff(sj, ps, d1, d2);
        const float a = z[j] * d1, b = et[j] * d2;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          L[k] = fmaf(V[k * K + j], sp, L[k]);
          nv[k] = fmaf(V[k * K + j], a, nv[k]);
          ne[k] = fmaf(V[k * K + j], b, ne[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
        xn[k] = __fadd_rn(__fmul_rn(sc.sqrt_atn, __fsub_rn(x0[k], L[k])), __fadd_rn(nv[k], ne[k]));
    }
    G.store(out1, img, xn);
  }
}

template <int K, int MODE>
static void local_launch(int fn, const float* in0, const float* in1, long long in1_stride, const float* in2, const float* y,
                         const float* V, float u00, float s0, const StepScalars& sc, float* out0, float* out1, long long groups,
                         int C, int D, cudaStream_t st) {
  const int grid = (int)cdivll(groups, 128);
#define LL(F) local_kernel<K, MODE, F><<<grid, 128, 0, st>>>(in0, in1, in1_stride, in2, y, V, u00, s0, sc, out0, out1, groups, C, D)
  switch (fn) {
    case LF_A: LL(LF_A); break;
    case LF_PINV: LL(LF_PINV); break;
    case LF_PROJECT: LL(LF_PROJECT); break;
    case LF_LAMBDA: LL(LF_LAMBDA); break;
    case LF_NOISE: LL(LF_NOISE); break;
    default: LL(LF_STEP); break;
  }
#undef LL
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// Inpainting (svd_operators.py:324-439; mask -> indices at diffusion.py:464-471).  y holds the kept entries of the
// (pixel, channel)-interleaved image in ascending order: y[rank(p*C + c)] = x[c][p].  Pure data movement: bit-exact.
// ------------------------------------------------------------------------------------------------------------------
template <int FN>
__global__ void inpaint_kernel(const float* __restrict__ in0, const float* __restrict__ in1, long long in1_stride,
                               const float* __restrict__ in2, const float* __restrict__ y, const int* __restrict__ rank,
                               StepScalars sc, float* __restrict__ out0, float* __restrict__ out1, int B, int C, int HW, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long img = (long long)C * HW;
  if (i >= (long long)B * img) return;
  const int p = (int)(i % HW);
  const int c = (int)((i / HW) % C);
  const int b = (int)(i / img);
  const int r = rank[(long long)p * C + c];
  const bool kept = r >= 0;
  const long long yi = (long long)b * M + r;
  const PlusScalars& ps = sc.plus;
  const float s = kept ? 1.f : 0.f;
  if (FN == LF_A) {
    if (kept) out0[yi] = in0[i];
  } else if (FN == LF_PINV) {
    out0[i] = kept ? y[yi] : 0.f;
  } else if (FN == LF_PROJECT) {
    const float x0 = in0[i];
    out0[i] = kept ? __fsub_rn(x0, __fsub_rn(x0, y[yi])) : __fsub_rn(x0, 0.f);
  } else if (FN == LF_LAMBDA) {
    out0[i] = __fmul_rn(in0[i], lam_coeff(s, ps));
  } else if (FN == LF_NOISE) {
    float d1, d2;
    noise_coeff(s, ps, d1, d2);
    out0[i] = __fadd_rn(__fmul_rn(in0[i], d1), __fmul_rn(in1[(long long)b * in1_stride + (i - (long long)b * img)], d2));
  } else {
    const float et = in1[(long long)b * in1_stride + (i - (long long)b * img)];
    const float z = in2[i];
    const float x0 = x0_from(in0[i], et, sc);
    out0[i] = x0;
    const float resid = kept ? __fsub_rn(x0, y[yi]) : 0.f;
    if (!sc.use_plus) {
      out1[i] = renoise(__fsub_rn(x0, resid), z, et, sc);
    } else {
      float d1, d2;
      noise_coeff(s, ps, d1, d2);
      const float x0h = __fsub_rn(x0, __fmul_rn(resid, lam_coeff(s, ps)));
      out1[i] = __fadd_rn(__fmul_rn(sc.sqrt_atn, x0h), __fadd_rn(__fmul_rn(z, d1), __fmul_rn(et, d2)));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Walsh-Hadamard (svd_operators.py:211-320).  H_{D*D} = H_D (x) H_D: butterflies over the low log2(D) index bits run
// along image rows, the high bits along columns; stage order equals the reference's h = 1, 2, 4, ... loop so every
// output is produced by the same sequence of fp32 adds.  The 1/D normalisation is applied after the last stage.
// ------------------------------------------------------------------------------------------------------------------
__global__ void fwht_rows_kernel(float* __restrict__ buf, int D, long long rows) {
  extern __shared__ float sm[];
  const int rpb = blockDim.x * 2 / D;  // rows per block (each thread owns 2 elements)
  const long long row0 = (long long)blockIdx.x * rpb;
  const int lr = (threadIdx.x * 2) / D;
  const long long row = row0 + lr;
  float* s = sm + lr * D;
  const int t = threadIdx.x % (D / 2);
  if (row < rows) {
    s[t] = buf[row * D + t];
    s[t + D / 2] = buf[row * D + t + D / 2];
  }
  __syncthreads();
  for (int h = 1; h < D; h <<= 1) {
    const int i = (t / h) * 2 * h + (t % h);
    const float a = s[i], b = s[i + h];
    __syncthreads();
    s[i] = a + b;
    s[i + h] = a - b;
    __syncthreads();
  }
  if (row < rows) {
    buf[row * D + t] = s[t];
    buf[row * D + t + D / 2] = s[t + D / 2];
  }
}
// columns: block = (image, 32-column strip); smem [D][33]
__global__ void fwht_cols_kernel(float* __restrict__ buf, int D, float scale) {
  extern __shared__ float sm[];
  const long long img = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  float* base = buf + img * D * D;
  const int cx = threadIdx.x % 32, ry = threadIdx.x / 32;
  const int rstep = blockDim.x / 32;
  for (int r = ry; r < D; r += rstep) sm[r * 33 + cx] = base[(long long)r * D + c0 + cx];
  __syncthreads();
  for (int h = 1; h < D; h <<= 1) {
    for (int t = ry; t < D / 2; t += rstep) {
      const int i = (t / h) * 2 * h + (t % h);
      const float a = sm[i * 33 + cx], b = sm[(i + h) * 33 + cx];
      sm[i * 33 + cx] = a + b;
      sm[(i + h) * 33 + cx] = a - b;
    }
    __syncthreads();
  }
  for (int r = ry; r < D; r += rstep) base[(long long)r * D + c0 + cx] = sm[r * 33 + cx] / scale;
}

// spectral-domain elementwise stage of the WH operator.  kept(c,q) <=> invperm[q]*C + c < M
template <int FN>
__global__ void wh_spec_kernel(const float* __restrict__ F, const float* __restrict__ F2, const float* __restrict__ y,
                               const int* __restrict__ perm, const int* __restrict__ invperm, PlusScalars ps,
                               float* __restrict__ out, int B, int C, int n2, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (FN == LF_A) {  // gather: out[b][j], j = p*C + c
    if (i >= (long long)B * M) return;
    const long long j = i % M;
    const int b = (int)(i / M);
    const int p = (int)(j / C), c = (int)(j % C);
    out[i] = F[((long long)b * C + c) * n2 + perm[p]];
    return;
  }
  if (i >= (long long)B * C * n2) return;
  const int q = (int)(i % n2);
  const int c = (int)((i / n2) % C);
  const int b = (int)(i / ((long long)C * n2));
  const long long j = (long long)invperm[q] * C + c;
  const bool kept = j < M;
  const float s = kept ? 1.f : 0.f;
  if (FN == LF_PINV) {
    out[i] = kept ? y[(long long)b * M + j] : 0.f;
  } else if (FN == LF_PROJECT) {
    out[i] = kept ? __fsub_rn(F[i], y[(long long)b * M + j]) : 0.f;
  } else if (FN == LF_LAMBDA) {
    out[i] = __fmul_rn(F[i], lam_coeff(s, ps));
  } else {  // LF_NOISE: raw pixels scaled in place of spectral coordinates
    float d1, d2;
    noise_coeff(s, ps, d1, d2);
    out[i] = __fadd_rn(__fmul_rn(F[i], d1), __fmul_rn(F2[i], d2));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise helpers of the generic (non-fused) step and of the separable operators
// ------------------------------------------------------------------------------------------------------------------
__global__ void x0_kernel(const float* __restrict__ xt, const float* __restrict__ et, long long et_stride, StepScalars sc,
                          float* __restrict__ x0, float* __restrict__ et3, int B, long long img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * img) return;
  const int b = (int)(i / img);
  const float e = et[(long long)b * et_stride + (i - (long long)b * img)];
  et3[i] = e;
  x0[i] = x0_from(xt[i], e, sc);
}
__global__ void final_ddnm_kernel(const float* __restrict__ x0, const float* __restrict__ resid, const float* __restrict__ z,
                                  const float* __restrict__ et3, StepScalars sc, float* __restrict__ xn, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xn[i] = renoise(__fsub_rn(x0[i], resid[i]), z[i], et3[i], sc);
}
__global__ void final_plus_kernel(const float* __restrict__ x0, const float* __restrict__ L, const float* __restrict__ nz,
                                  StepScalars sc, float* __restrict__ xn, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xn[i] = __fadd_rn(__fmul_rn(sc.sqrt_atn, __fsub_rn(x0[i], L[i])), nz[i]);
}
__global__ void sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = __fsub_rn(a[i], b[i]);
}
// t[b][c][pos] *= tab[(per_channel ? c : 0)][pos]
__global__ void mul_table_kernel(float* __restrict__ t, const float* __restrict__ tab, int per_channel, int C, int n2, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = (int)(i % n2);
  const int c = (int)((i / n2) % C);
  t[i] = __fmul_rn(t[i], tab[(per_channel ? (long long)c * n2 : 0) + pos]);
}
// deblur Lambda tables from the un-thresholded singulars at each spectral position
template <int FN>
__global__ void deblur_coeff_kernel(const float* __restrict__ v, const float* __restrict__ e, const float* __restrict__ sorig,
                                    PlusScalars ps, float* __restrict__ out, int n2, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sorig[i % n2];
  if (FN == LF_LAMBDA) {
    out[i] = __fmul_rn(v[i], lam_coeff(s, ps));
  } else {
    float d1, d2;
    noise_coeff(s, ps, d1, d2);
    out[i] = __fadd_rn(__fmul_rn(v[i], d1), __fmul_rn(e[i], d2));
  }
}

// CS (svd_operators.py:101-159): 32x32 patches <-> rows of a [B*C*y*y, 1024] matrix (row-major inside the patch)
template <bool TO_ROWS>
__global__ void cs_patch_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * D * D;
  if (i >= total) return;
  const int yd = D / 32;
  const int k = (int)(i % 1024);
  const long long pr = i / 1024;                 // (b, c, py, px)
  const int px = (int)(pr % yd), py = (int)((pr / yd) % yd);
  const long long bc = pr / ((long long)yd * yd);
  const long long img = bc * D * D + (long long)(py * 32 + k / 32) * D + (px * 32 + k % 32);
  if (TO_ROWS) dst[i] = src[img];
  else dst[img] = src[i];
}

// Denoising (svd_operators.py:442-476): A = I; Lambda / Lambda_noise are SCALAR rules of their own (not the table rule)
template <int FN>
__global__ void denoise_kernel(const float* __restrict__ in0, const float* __restrict__ in1, long long in1_stride,
                               const float* __restrict__ in2, const float* __restrict__ y, StepScalars sc, float* __restrict__ out0,
                               float* __restrict__ out1, int B, long long img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * img) return;
  const int b = (int)(i / img);
  const PlusScalars& ps = sc.plus;
  const float asy = __fmul_rn(ps.a, ps.sigma_y);
  // Lambda (:462-467): sigma_t < a*sigma_y ? v * (sigma_t * sqrt(1-eta^2) / a / sigma_y) : v
  const float lam = (ps.sigma_t < asy) ? __fdiv_rn(__fdiv_rn(__fmul_rn(ps.sigma_t, ps.c), ps.a), ps.sigma_y) : 1.0f;
  // Lambda_noise (:469-474): sigma_t >= a*sigma_y ? v * sqrt(sigma_t^2 - a^2 sigma_y^2) : v * sigma_t * eta   (epsilon unused)
  const float t2 = __fsub_rn(__fmul_rn(ps.sigma_t, ps.sigma_t), __fmul_rn(__fmul_rn(ps.a, ps.a), ps.sy2));
  if (FN == LF_LAMBDA) {
    out0[i] = (ps.sigma_t < asy) ? __fmul_rn(in0[i], lam) : in0[i];
  } else if (FN == LF_NOISE) {
    out0[i] = (ps.sigma_t >= asy) ? __fmul_rn(in0[i], sqrtf(t2)) : __fmul_rn(__fmul_rn(in0[i], ps.sigma_t), ps.eta);
  } else {  // LF_STEP
    const float et = in1[(long long)b * in1_stride + (i - (long long)b * img)];
    const float z = in2[i];
    const float x0 = x0_from(in0[i], et, sc);
    out0[i] = x0;
    const float resid = __fsub_rn(x0, y[i]);
    if (!sc.use_plus) {
      out1[i] = renoise(__fsub_rn(x0, resid), z, et, sc);
    } else {
      const float L = (ps.sigma_t < asy) ? __fmul_rn(resid, lam) : resid;
      const float nz = (ps.sigma_t >= asy) ? __fmul_rn(z, sqrtf(t2)) : __fmul_rn(__fmul_rn(z, ps.sigma_t), ps.eta);
      out1[i] = __fadd_rn(__fmul_rn(sc.sqrt_atn, __fsub_rn(x0, L)), nz);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Operator
// ------------------------------------------------------------------------------------------------------------------
template <class T>
static T* upload(std::vector<void*>& owned, const T* host, size_t n) {
  T* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
  CUDA_CHECK(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
  owned.push_back(d);
  return d;
}
static std::vector<float> transpose(const float* m, int r, int c) {
  std::vector<float> t((size_t)r * c);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) t[(size_t)j * r + i] = m[(size_t)i * c + j];
  return t;
}

Operator::Operator(int kind, int channels, int img_dim, int ratio, const float* v_small, const float* u_small,
                   const float* singulars, const float* singulars_orig, const long long* perm, const long long* mask,
                   const float* v_small2, const float* u_small2)
    : kind_(kind), C_(channels), D_(img_dim), ratio_(ratio) {
  const int n2 = D_ * D_;
  DDNM_CHECK(channels >= 1 && img_dim >= 2, "bad operator geometry");
  N_ = (long long)C_ * n2;
  switch (kind) {
    case OP_SR: {
      DDNM_CHECK(ratio >= 1 && img_dim % ratio == 0, "img_dim % ratio");  // svd_operators.py:481
      // 2 / 4 / 8: one thread per patch with the basis in shared memory; any other ratio (evaluation.sh runs 16x): patch rows +
      // the K x K basis as a small GEMM
      sr_generic_ = !(ratio == 2 || ratio == 4 || ratio == 8);
      DDNM_CHECK(ratio <= 64, "SuperResolution: ratio <= 64");
      DDNM_CHECK(v_small && u_small && singulars, "SuperResolution needs V_small, U_small, singulars_small");
      const int K = ratio * ratio;
      V_ = upload(owned_, v_small, (size_t)K * K);
      if (sr_generic_) {
        std::vector<float> v0(K);
        for (int k = 0; k < K; ++k) v0[k] = v_small[(size_t)k * K];
        v0_ = upload(owned_, v0.data(), (size_t)K);
      }
      u00_ = u_small[0];
      s0_ = singulars[0];
      M_ = (long long)C_ * (D_ / ratio) * (D_ / ratio);
      break;
    }
    case OP_COLOR: {
      DDNM_CHECK(channels == 3 && v_small && u_small && singulars, "Colorization needs 3 channels and its 3x3 basis");
      V_ = upload(owned_, v_small, 9);
      u00_ = u_small[0];
      s0_ = singulars[0];
      M_ = n2;
      break;
    }
    case OP_INPAINT: {
      DDNM_CHECK(mask, "Inpainting needs the mask");
      const long long ne = (long long)n2 * C_;
      std::vector<int> rank(ne);
      int r = 0;
      for (long long e = 0; e < ne; ++e) rank[e] = mask[e] != 0 ? r++ : -1;
      rank_ = upload(owned_, rank.data(), ne);
      M_ = r;
      break;
    }
    case OP_WH: {
      DDNM_CHECK(perm && ratio >= 1, "WalshHadamardCS needs perm");
      DDNM_CHECK((D_ & (D_ - 1)) == 0 && D_ >= 32 && D_ <= 1024, "WalshHadamardCS: img_dim must be a power of two in [32, 1024]");
      std::vector<int> p(n2), ip(n2, -1);
      for (int i = 0; i < n2; ++i) {
        DDNM_CHECK(perm[i] >= 0 && perm[i] < n2 && ip[perm[i]] < 0, "perm is not a permutation");
        p[i] = (int)perm[i];
        ip[perm[i]] = i;
      }
      perm_ = upload(owned_, p.data(), n2);
      invperm_ = upload(owned_, ip.data(), n2);
      M_ = (long long)C_ * n2 / ratio;
      break;
    }
    case OP_DENOISE:
      M_ = (long long)C_ * n2;   // svd_operators.py:442-476: A = identity
      break;
    case OP_CS: {
      // ratio field carries cs_size = int(32*32*cs_ratio) (svd_operators.py:111)
      DDNM_CHECK(v_small && img_dim % 32 == 0 && ratio >= 1 && ratio <= 1024, "CS needs V_small [1024,1024], img_dim % 32, 1 <= cs_size <= 1024");
      V_ = upload(owned_, v_small, (size_t)1024 * 1024);
      cs_size_ = ratio;
      M_ = (long long)C_ * (D_ / 32) * (D_ / 32) * cs_size_;
      break;
    }
    case OP_DEBLUR:
    case OP_DEBLUR2D: {
      if (kind == OP_DEBLUR) DDNM_CHECK(singulars_orig != nullptr, "Deblurring needs the un-thresholded singulars");
      DDNM_CHECK(v_small && u_small && singulars && perm, "Deblurring needs U, V, singular tables and perm");
      V_ = upload(owned_, v_small, (size_t)n2);
      U_ = upload(owned_, u_small, (size_t)n2);
      auto vt = transpose(v_small, D_, D_), ut = transpose(u_small, D_, D_);
      Vt_ = upload(owned_, vt.data(), (size_t)n2);
      Ut_ = upload(owned_, ut.data(), (size_t)n2);
      Vr_ = V_; Vrt_ = Vt_; Ur_ = U_; Urt_ = Ut_;
      if (kind == OP_DEBLUR2D) {   // svd_operators.py:1094-1166: different 1-D factors on the two sides, no Lambda
        DDNM_CHECK(v_small2 && u_small2, "Deblurring2D needs the second pair of factors");
        auto vt2 = transpose(v_small2, D_, D_), ut2 = transpose(u_small2, D_, D_);
        Vr_ = upload(owned_, v_small2, (size_t)n2);
        Ur_ = upload(owned_, u_small2, (size_t)n2);
        Vrt_ = upload(owned_, vt2.data(), (size_t)n2);
        Urt_ = upload(owned_, ut2.data(), (size_t)n2);
      }
      // singulars() = _singulars.repeat(1, 3) is TILED while spectral vectors are (pos, chan)-interleaved
      // (svd_operators.py:1001 vs :984): D[c][perm[p]] = S[(C*p + c) mod n2]
      std::vector<float> tD((size_t)C_ * n2), tDi((size_t)C_ * n2), tS(n2);
      for (int p = 0; p < n2; ++p) {
        const long long q = perm[p];
        DDNM_CHECK(q >= 0 && q < n2, "bad perm entry");
        for (int c = 0; c < C_; ++c) {
          const float s = singulars[((long long)C_ * p + c) % n2];
          tD[(size_t)c * n2 + q] = s;
          tDi[(size_t)c * n2 + q] = s == 0.f ? 0.f : 1.0f / s;
        }
        tS[q] = singulars_orig ? singulars_orig[p] : 0.f;
      }
      tabD_ = upload(owned_, tD.data(), tD.size());
      tabDinv_ = upload(owned_, tDi.data(), tDi.size());
      tabSorig_ = upload(owned_, tS.data(), tS.size());
      M_ = (long long)C_ * n2;
      break;
    }
    case OP_SRCONV: {
      DDNM_CHECK(v_small && u_small && singulars && ratio >= 1 && img_dim % ratio == 0, "SRConv needs U_small, V_small, singulars_small");
      const int sm = D_ / ratio;
      std::vector<float> vk((size_t)D_ * sm);
      for (int i = 0; i < D_; ++i)
        for (int j = 0; j < sm; ++j) vk[(size_t)i * sm + j] = v_small[(size_t)i * D_ + j];
      auto vkt = transpose(vk.data(), D_, sm);
      auto ut = transpose(u_small, sm, sm);
      V_ = upload(owned_, vk.data(), vk.size());     // D x sm
      Vt_ = upload(owned_, vkt.data(), vkt.size());  // sm x D
      U_ = upload(owned_, u_small, (size_t)sm * sm);
      Ut_ = upload(owned_, ut.data(), ut.size());
      std::vector<float> s2((size_t)sm * sm), s2i((size_t)sm * sm);
      for (int i = 0; i < sm; ++i)
        for (int j = 0; j < sm; ++j) {
          const float s = singulars[i] * singulars[j];
          s2[(size_t)i * sm + j] = s;
          s2i[(size_t)i * sm + j] = s == 0.f ? 0.f : 1.0f / s;
        }
      tabD_ = upload(owned_, s2.data(), s2.size());
      tabDinv_ = upload(owned_, s2i.data(), s2i.size());
      M_ = (long long)C_ * sm * sm;
      break;
    }
    case OP_GENERAL: {
      // GeneralA (svd_operators.py:173-208): dense A = U diag(s) V^T with FULL factors U [m,m], V [n,n]; only the first m
      // columns of V ever meet a non-zero coefficient (A(): temp[:, :m]; A_pinv(): add_zeros pads m..n with zeros).
      // geometry: x is a flat vector of n = img_dim entries per row (channels must be 1)
      DDNM_CHECK(C_ == 1, "GeneralA: pass channels = 1 and img_dim = n (columns of A)");
      const long long n = N_ = D_;
      const int m = ratio;
      DDNM_CHECK(v_small && u_small && singulars && m >= 1 && m <= n, "GeneralA needs U [m,m], V [n,n], singulars [m] with m <= n");
      std::vector<float> vm((size_t)n * m), sinv(m);
      for (long long i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) vm[(size_t)i * m + j] = v_small[(size_t)i * n + j];
      for (int j = 0; j < m; ++j) sinv[j] = singulars[j] == 0.f ? 0.f : 1.0f / singulars[j];   // A_pinv's factors (:74-75)
      V_ = upload(owned_, vm.data(), vm.size());          // n x m
      U_ = upload(owned_, u_small, (size_t)m * m);
      tabD_ = upload(owned_, singulars, (size_t)m);
      tabDinv_ = upload(owned_, sinv.data(), (size_t)m);
      M_ = m;
      break;
    }
    default:
      throw Error("unknown operator kind");
  }
}

Operator::~Operator() {
  for (void* p : owned_) cudaFree(p);
  for (float* p : scr_)
    if (p) cudaFree(p);
}

float* Operator::scratch(int idx, size_t elems) {
  if (scr_elems_[idx] < elems) {
    if (scr_[idx]) {
      CUDA_CHECK(cudaDeviceSynchronize());
      CUDA_CHECK(cudaFree(scr_[idx]));
    }
    CUDA_CHECK(cudaMalloc(&scr_[idx], elems * sizeof(float)));
    scr_elems_[idx] = elems;
  }
  return scr_[idx];
}

static inline int blocks(long long n, int t = 256) { return (int)cdivll(n, t); }

template <int FN>
static void local_dispatch(int kind, int ratio, const float* in0, const float* in1, long long in1_stride, const float* in2,
                           const float* y, const float* V, float u00, float s0, const StepScalars& sc, float* out0, float* out1,
                           int B, int C, int D, cudaStream_t st) {
  if (kind == OP_COLOR) {
    local_launch<3, 1>(FN, in0, in1, in1_stride, in2, y, V, u00, s0, sc, out0, out1, (long long)B * D * D, C, D, st);
  } else {
    const long long groups = (long long)B * C * (D / ratio) * (D / ratio);
    if (ratio == 2) local_launch<4, 0>(FN, in0, in1, in1_stride, in2, y, V, u00, s0, sc, out0, out1, groups, C, D, st);
    else if (ratio == 4) local_launch<16, 0>(FN, in0, in1, in1_stride, in2, y, V, u00, s0, sc, out0, out1, groups, C, D, st);
    else local_launch<64, 0>(FN, in0, in1, in1_stride, in2, y, V, u00, s0, sc, out0, out1, groups, C, D, st);
  }
}

void Operator::fwht(float* buf, int B, cudaStream_t s) {
  const long long rows = (long long)B * C_ * D_;
  const int threads = std::max(128, D_ / 2);
  const int rpb = threads * 2 / D_;
  fwht_rows_kernel<<<(int)cdivll(rows, rpb), threads, (size_t)rpb * D_ * 4, s>>>(buf, D_, rows);
  CUDA_CHECK(cudaGetLastError());
  const size_t smem = (size_t)D_ * 33 * 4;
  // the attribute is per device: set it whenever more than the default 48 KiB is needed (a cached process-wide flag would leave
  // a second GPU of the same process without it)
  if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(fwht_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(D_ / 32, B * C_);
  fwht_cols_kernel<<<grid, 256, smem, s>>>(buf, D_, (float)D_);
  CUDA_CHECK(cudaGetLastError());
}

// out[b] = L (lr x lc) * X[b] (lc x rr) * R (rr x rc), X: B*C images; T: scratch of lr*rr per image
void Operator::sandwich(const float* L, int lr, int lc, const float* X, int B, const float* R, int rr, int rc, float* T,
                        float* out, cudaStream_t s) {
  const int nb = B * C_;
  sgemm_batched(false, nb, 1, lr, rr, lc, 1.0f, L, lc, 0, 0, X, rr, (long long)lc * rr, 0, T, rr, (long long)lr * rr, 0, s);
  sgemm_batched(false, nb, 1, lr, rc, rr, 1.0f, T, rr, (long long)lr * rr, 0, R, rc, 0, 0, out, rc, (long long)lr * rc, 0, s);
}

void Operator::deblur_A(const float* x, int B, float* y, cudaStream_t s) {
  if (kind_ == OP_CS) return cs_A(x, B, y, s);
  if (kind_ == OP_GENERAL) return general_A(x, B, y, s);
  const int n2 = D_ * D_;
  const long long n = (long long)B * C_ * n2;
  if (kind_ == OP_DEBLUR || kind_ == OP_DEBLUR2D) {
    float* T = scratch(0, n);
    float* S = scratch(1, n);
    sandwich(Vt_, D_, D_, x, B, Vr_, D_, D_, T, S, s);
    mul_table_kernel<<<blocks(n), 256, 0, s>>>(S, tabD_, 1, C_, n2, n);
    sandwich(U_, D_, D_, S, B, Urt_, D_, D_, T, y, s);
  } else {  // SRConv: Vk^T X Vk -> (sm x sm), scale, U . U^T
    const int sm = D_ / ratio_;
    float* T = scratch(0, (size_t)B * C_ * sm * D_);
    float* S = scratch(1, (size_t)B * C_ * sm * sm);
    float* T2 = scratch(2, (size_t)B * C_ * sm * sm);
    sandwich(Vt_, sm, D_, x, B, V_, D_, sm, T, S, s);
    const long long ns = (long long)B * C_ * sm * sm;
    mul_table_kernel<<<blocks(ns), 256, 0, s>>>(S, tabD_, 0, C_, sm * sm, ns);
    sandwich(U_, sm, sm, S, B, Ut_, sm, sm, T2, y, s);
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::deblur_Apinv(const float* y, int B, float* x, cudaStream_t s) {
  if (kind_ == OP_CS) return cs_Apinv(y, B, x, s);
  if (kind_ == OP_GENERAL) return general_Apinv(y, B, x, s);
  const int n2 = D_ * D_;
  const long long n = (long long)B * C_ * n2;
  if (kind_ == OP_DEBLUR || kind_ == OP_DEBLUR2D) {
    float* T = scratch(0, n);
    float* S = scratch(1, n);
    sandwich(Ut_, D_, D_, y, B, Ur_, D_, D_, T, S, s);
    mul_table_kernel<<<blocks(n), 256, 0, s>>>(S, tabDinv_, 1, C_, n2, n);
    sandwich(V_, D_, D_, S, B, Vrt_, D_, D_, T, x, s);
  } else {
    const int sm = D_ / ratio_;
    float* S = scratch(1, (size_t)B * C_ * sm * sm);
    float* T2 = scratch(2, (size_t)B * C_ * sm * sm);
    float* T = scratch(0, (size_t)B * C_ * sm * D_);
    sandwich(Ut_, sm, sm, y, B, U_, sm, sm, T2, S, s);
    const long long ns = (long long)B * C_ * sm * sm;
    mul_table_kernel<<<blocks(ns), 256, 0, s>>>(S, tabDinv_, 0, C_, sm * sm, ns);
    // x = Vk (D x sm) * S (sm x sm) * Vk^T (sm x D): first product is D x sm per image
    const int nb = B * C_;
    sgemm_batched(false, nb, 1, D_, sm, sm, 1.0f, V_, sm, 0, 0, S, sm, (long long)sm * sm, 0, T, sm, (long long)D_ * sm, 0, s);
    sgemm_batched(false, nb, 1, D_, D_, sm, 1.0f, T, sm, (long long)D_ * sm, 0, Vt_, D_, 0, 0, x, D_, (long long)D_ * D_, 0, s);
  }
  CUDA_CHECK(cudaGetLastError());
}

// GeneralA: y = U (s * (V^T x)[:m]),  x = V add_zeros((1/s) * (U^T y))  — three small fp32 GEMMs over the batch
void Operator::general_A(const float* x, int B, float* y, cudaStream_t s) {
  const int n = (int)x_dim(), m = (int)M_;
  float* T = scratch(0, (size_t)B * m);
  sgemm_batched(false, 1, 1, B, m, n, 1.0f, x, n, 0, 0, V_, m, 0, 0, T, m, 0, 0, s);          // T = X Vm
  mul_table_kernel<<<blocks((long long)B * m), 256, 0, s>>>(T, tabD_, 0, 1, m, (long long)B * m);
  sgemm_batched(true, 1, 1, B, m, m, 1.0f, T, m, 0, 0, U_, m, 0, 0, y, m, 0, 0, s);           // y[b][k] = sum_j U[k][j] T[b][j]
}
void Operator::general_Apinv(const float* y, int B, float* x, cudaStream_t s) {
  const int n = (int)x_dim(), m = (int)M_;
  float* T = scratch(0, (size_t)B * m);
  sgemm_batched(false, 1, 1, B, m, m, 1.0f, y, m, 0, 0, U_, m, 0, 0, T, m, 0, 0, s);          // T[b][j] = sum_k y[b][k] U[k][j]
  mul_table_kernel<<<blocks((long long)B * m), 256, 0, s>>>(T, tabDinv_, 0, 1, m, (long long)B * m);
  sgemm_batched(true, 1, 1, B, n, m, 1.0f, T, m, 0, 0, V_, m, 0, 0, x, n, 0, 0, s);           // x[b][i] = sum_j V[i][j] T[b][j]
}

void Operator::cs_A(const float* x, int B, float* y, cudaStream_t s) {
  const long long n = (long long)B * C_ * D_ * D_;
  const int rows = (int)(n / 1024);
  float* P = scratch(0, n);
  cs_patch_kernel<true><<<blocks(n), 256, 0, s>>>(x, P, B, C_, D_);
  // first cs_size coefficients of V^T patch  ==  P [rows x 1024] . V[:, :cs]
  sgemm_batched(false, 1, 1, rows, cs_size_, 1024, 1.0f, P, 1024, 0, 0, V_, 1024, 0, 0, y, cs_size_, 0, 0, s);
}
void Operator::cs_Apinv(const float* y, int B, float* x, cudaStream_t s) {
  const long long n = (long long)B * C_ * D_ * D_;
  const int rows = (int)(n / 1024);
  float* P = scratch(0, n);
  // V (c, 0, ..)  ==  Y [rows x cs] . V[:, :cs]^T
  sgemm_batched(true, 1, 1, rows, 1024, cs_size_, 1.0f, y, cs_size_, 0, 0, V_, 1024, 0, 0, P, 1024, 0, 0, s);
  cs_patch_kernel<false><<<blocks(n), 256, 0, s>>>(P, x, B, C_, D_);
}

// ------------------------------------------------------------------------------------------------------------------
// SuperResolution, generic ratio r (K = r*r entries per patch; svd_operators.py:479-623 with r = 16 in evaluation.sh):
// patches become rows of a [B*C*y*y, K] matrix (row-major inside the patch, the reference's unfold order :510-512).  A has rank 1
// per patch, so A / A^+ / the projection / Lambda need only V[:, 0] and one dot product per row (V is orthogonal:
// V diag(l0, lz, .., lz) V^T x = lz x + (l0 - lz) <v0, x> v0); Lambda_noise multiplies RAW pixels by V (:581-621), which is a
// [rows, K] x V^T GEMM on the CUDA cores.
// ------------------------------------------------------------------------------------------------------------------
template <bool TO_ROWS>
__global__ void sr_patch_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int D, int r) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * D * D;
  if (i >= total) return;
  const int K = r * r, yd = D / r;
  const int k = (int)(i % K);
  const long long pr = i / K;                 // (b, c, py, px)
  const int px = (int)(pr % yd), py = (int)((pr / yd) % yd);
  const long long bc = pr / ((long long)yd * yd);
  const long long img = bc * D * D + (long long)(py * r + k / r) * D + (px * r + k % r);
  if (TO_ROWS) dst[i] = src[img];
  else dst[img] = src[i];
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// one warp per patch row.  P: rows [rows, K]; Q / Wv / We: further row operands; yv: [rows]
template <int FN>
__global__ void srg_rows_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ Wv,
                                const float* __restrict__ We, const float* __restrict__ yv, const float* __restrict__ v0, float u00,
                                float s0, PlusScalars ps, float* __restrict__ out, long long rows, int K) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = P ? P + row * K : nullptr;
  float dot = 0.f;
  if (FN == LF_A || FN == LF_PROJECT || FN == LF_LAMBDA) {
    for (int k = lane; k < K; k += 32) dot = fmaf(v0[k], p[k], dot);
    dot = warp_sum(dot);
  }
  if (FN == LF_A) {
    if (lane == 0) out[row] = __fmul_rn(u00, __fmul_rn(s0, dot));
  } else if (FN == LF_PINV) {
    const float cc = __fmul_rn(__fmul_rn(u00, yv[row]), __fdiv_rn(1.0f, s0));
    for (int k = lane; k < K; k += 32) out[row * K + k] = __fmul_rn(v0[k], cc);
  } else if (FN == LF_PROJECT) {
    const float r = __fsub_rn(__fmul_rn(u00, __fmul_rn(s0, dot)), yv[row]);
    const float cc = __fmul_rn(__fmul_rn(u00, r), __fdiv_rn(1.0f, s0));
    for (int k = lane; k < K; k += 32) out[row * K + k] = __fsub_rn(p[k], __fmul_rn(v0[k], cc));
  } else if (FN == LF_LAMBDA) {
    const float l0 = lam_coeff(s0, ps), lz = lam_coeff(0.f, ps);
    const float t = __fmul_rn(__fsub_rn(l0, lz), dot);
    for (int k = lane; k < K; k += 32) out[row * K + k] = fmaf(v0[k], t, __fmul_rn(lz, p[k]));
  } else {  // LF_NOISE: P = raw v rows, Q = raw eps rows, Wv = P V^T, We = Q V^T
    float d10, d20, d1z, d2z;
    noise_coeff(s0, ps, d10, d20);
    noise_coeff(0.f, ps, d1z, d2z);
    const float a0 = __fmul_rn(__fsub_rn(d10, d1z), p[0]);
    const float b0 = __fmul_rn(__fsub_rn(d20, d2z), Q[row * K]);
    for (int k = lane; k < K; k += 32) {
      const float ov = fmaf(v0[k], a0, __fmul_rn(d1z, Wv[row * K + k]));
      const float oe = fmaf(v0[k], b0, __fmul_rn(d2z, We[row * K + k]));
      out[row * K + k] = __fadd_rn(ov, oe);
    }
  }
}

void Operator::srg_A(const float* x, int B, float* y, cudaStream_t s) {
  const long long n = (long long)B * N_;
  const int K = ratio_ * ratio_;
  const long long rows = n / K;
  float* P = scratch(6, n);
  sr_patch_kernel<true><<<blocks(n), 256, 0, s>>>(x, P, B, C_, D_, ratio_);
  srg_rows_kernel<LF_A><<<blocks(rows * 32), 256, 0, s>>>(P, nullptr, nullptr, nullptr, nullptr, v0_, u00_, s0_, PlusScalars{}, y, rows, K);
}
void Operator::srg_Apinv(const float* y, int B, float* x, cudaStream_t s) {
  const long long n = (long long)B * N_;
  const int K = ratio_ * ratio_;
  const long long rows = n / K;
  float* P = scratch(6, n);
  srg_rows_kernel<LF_PINV><<<blocks(rows * 32), 256, 0, s>>>(nullptr, nullptr, nullptr, nullptr, y, v0_, u00_, s0_, PlusScalars{}, P, rows, K);
  sr_patch_kernel<false><<<blocks(n), 256, 0, s>>>(P, x, B, C_, D_, ratio_);
}
void Operator::srg_project(const float* x0, const float* y, int B, float* out, cudaStream_t s) {
  const long long n = (long long)B * N_;
  const int K = ratio_ * ratio_;
  const long long rows = n / K;
  float* P = scratch(6, n);
  float* O = scratch(7, n);
  sr_patch_kernel<true><<<blocks(n), 256, 0, s>>>(x0, P, B, C_, D_, ratio_);
  srg_rows_kernel<LF_PROJECT><<<blocks(rows * 32), 256, 0, s>>>(P, nullptr, nullptr, nullptr, y, v0_, u00_, s0_, PlusScalars{}, O, rows, K);
  sr_patch_kernel<false><<<blocks(n), 256, 0, s>>>(O, out, B, C_, D_, ratio_);
}
void Operator::srg_lambda(const float* v, int B, const PlusScalars& ps, float* out, cudaStream_t s) {
  const long long n = (long long)B * N_;
  const int K = ratio_ * ratio_;
  const long long rows = n / K;
  float* P = scratch(6, n);
  float* O = scratch(7, n);
  sr_patch_kernel<true><<<blocks(n), 256, 0, s>>>(v, P, B, C_, D_, ratio_);
  srg_rows_kernel<LF_LAMBDA><<<blocks(rows * 32), 256, 0, s>>>(P, nullptr, nullptr, nullptr, nullptr, v0_, u00_, s0_, ps, O, rows, K);
  sr_patch_kernel<false><<<blocks(n), 256, 0, s>>>(O, out, B, C_, D_, ratio_);
}
void Operator::srg_lambda_noise(const float* v, const float* eps, int B, const PlusScalars& ps, float* out, cudaStream_t s) {
  const long long n = (long long)B * N_;
  const int K = ratio_ * ratio_;
  const long long rows = n / K;
  float* P = scratch(6, n);
  float* Q = scratch(7, n);
  float* Wv = scratch(8, n);
  float* We = scratch(9, n);
  sr_patch_kernel<true><<<blocks(n), 256, 0, s>>>(v, P, B, C_, D_, ratio_);
  sr_patch_kernel<true><<<blocks(n), 256, 0, s>>>(eps, Q, B, C_, D_, ratio_);
  // W[row][k] = sum_j V[k][j] * raw[row][j]
  sgemm_batched(true, 1, 1, (int)rows, K, K, 1.0f, P, K, 0, 0, V_, K, 0, 0, Wv, K, 0, 0, s);
  sgemm_batched(true, 1, 1, (int)rows, K, K, 1.0f, Q, K, 0, 0, V_, K, 0, 0, We, K, 0, 0, s);
  srg_rows_kernel<LF_NOISE><<<blocks(rows * 32), 256, 0, s>>>(P, Q, Wv, We, nullptr, v0_, u00_, s0_, ps, P, rows, K);
  sr_patch_kernel<false><<<blocks(n), 256, 0, s>>>(P, out, B, C_, D_, ratio_);
}

void Operator::A(const float* x, int B, float* y, cudaStream_t s) {
  StepScalars sc{};
  const int n2 = D_ * D_;
  if (kind_ == OP_CS) {
    cs_A(x, B, y, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (kind_ == OP_DENOISE) {
    CUDA_CHECK(cudaMemcpyAsync(y, x, (size_t)B * C_ * n2 * 4, cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (kind_ == OP_SR && sr_generic_) {
    srg_A(x, B, y, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  switch (kind_) {
    case OP_SR: case OP_COLOR:
      local_dispatch<LF_A>(kind_, ratio_, x, nullptr, 0, nullptr, nullptr, V_, u00_, s0_, sc, y, nullptr, B, C_, D_, s);
      break;
    case OP_INPAINT:
      inpaint_kernel<LF_A><<<blocks((long long)B * C_ * n2), 256, 0, s>>>(x, nullptr, 0, nullptr, nullptr, rank_, sc, y, nullptr, B, C_, n2, M_);
      break;
    case OP_WH: {
      const long long n = (long long)B * C_ * n2;
      float* F = scratch(0, n);
      CUDA_CHECK(cudaMemcpyAsync(F, x, n * 4, cudaMemcpyDeviceToDevice, s));
      fwht(F, B, s);
      wh_spec_kernel<LF_A><<<blocks((long long)B * M_), 256, 0, s>>>(F, nullptr, nullptr, perm_, invperm_, sc.plus, y, B, C_, n2, M_);
      break;
    }
    default: deblur_A(x, B, y, s);
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::A_pinv(const float* y, int B, float* x, cudaStream_t s) {
  StepScalars sc{};
  const int n2 = D_ * D_;
  if (kind_ == OP_CS) {
    cs_Apinv(y, B, x, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (kind_ == OP_DENOISE) {
    CUDA_CHECK(cudaMemcpyAsync(x, y, (size_t)B * C_ * n2 * 4, cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (kind_ == OP_SR && sr_generic_) {
    srg_Apinv(y, B, x, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  switch (kind_) {
    case OP_SR: case OP_COLOR:
      local_dispatch<LF_PINV>(kind_, ratio_, nullptr, nullptr, 0, nullptr, y, V_, u00_, s0_, sc, x, nullptr, B, C_, D_, s);
      break;
    case OP_INPAINT:
      inpaint_kernel<LF_PINV><<<blocks((long long)B * C_ * n2), 256, 0, s>>>(nullptr, nullptr, 0, nullptr, y, rank_, sc, x, nullptr, B, C_, n2, M_);
      break;
    case OP_WH:
      wh_spec_kernel<LF_PINV><<<blocks((long long)B * C_ * n2), 256, 0, s>>>(nullptr, nullptr, y, perm_, invperm_, sc.plus, x, B, C_, n2, M_);
      fwht(x, B, s);
      break;
    default: deblur_Apinv(y, B, x, s);
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::project(const float* x0, const float* y, int B, float* out, cudaStream_t s) {
  StepScalars sc{};
  const int n2 = D_ * D_;
  const long long n = (long long)B * N_;
  if (kind_ == OP_DENOISE) {   // x0 - (x0 - y)
    float* R = scratch(4, n);
    sub_kernel<<<blocks(n), 256, 0, s>>>(x0, y, R, n);
    sub_kernel<<<blocks(n), 256, 0, s>>>(x0, R, out, n);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (kind_ == OP_SR && sr_generic_) {
    srg_project(x0, y, B, out, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  switch (kind_) {
    case OP_SR: case OP_COLOR:
      local_dispatch<LF_PROJECT>(kind_, ratio_, x0, nullptr, 0, nullptr, y, V_, u00_, s0_, sc, out, nullptr, B, C_, D_, s);
      break;
    case OP_INPAINT:
      inpaint_kernel<LF_PROJECT><<<blocks(n), 256, 0, s>>>(x0, nullptr, 0, nullptr, y, rank_, sc, out, nullptr, B, C_, n2, M_);
      break;
    case OP_WH: {
      float* F = scratch(0, n);
      CUDA_CHECK(cudaMemcpyAsync(F, x0, n * 4, cudaMemcpyDeviceToDevice, s));
      fwht(F, B, s);
      float* R = scratch(1, n);
      wh_spec_kernel<LF_PROJECT><<<blocks(n), 256, 0, s>>>(F, nullptr, y, perm_, invperm_, sc.plus, R, B, C_, n2, M_);
      fwht(R, B, s);
      sub_kernel<<<blocks(n), 256, 0, s>>>(x0, R, out, n);
      break;
    }
    default: {
      float* Ay = scratch(3, (size_t)B * M_);
      float* R = scratch(4, n);
      deblur_A(x0, B, Ay, s);
      sub_kernel<<<blocks((long long)B * M_), 256, 0, s>>>(Ay, y, Ay, (long long)B * M_);
      deblur_Apinv(Ay, B, R, s);
      sub_kernel<<<blocks(n), 256, 0, s>>>(x0, R, out, n);
    }
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::lambda(const float* v, int B, const PlusScalars& ps, float* out, cudaStream_t s) {
  StepScalars sc{};
  sc.plus = ps;
  const int n2 = D_ * D_;
  const long long n = (long long)B * C_ * n2;
  if (kind_ == OP_DENOISE) {
    denoise_kernel<LF_LAMBDA><<<blocks(n), 256, 0, s>>>(v, nullptr, 0, nullptr, nullptr, sc, out, nullptr, B, (long long)C_ * n2);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (kind_ == OP_DEBLUR2D) throw Error("Deblurring2D defines no Lambda (svd_operators.py:1094-1166): sigma_y > 0 is unsupported, as in the reference");
  if (kind_ == OP_CS) throw Error("CS defines no Lambda (svd_operators.py:101-159): sigma_y > 0 is unsupported, as in the reference");
  if (kind_ == OP_GENERAL) throw Error("GeneralA defines no Lambda (svd_operators.py:173-208): sigma_y > 0 is unsupported, as in the reference");
  if (kind_ == OP_SR && sr_generic_) {
    srg_lambda(v, B, ps, out, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  switch (kind_) {
    case OP_SR: case OP_COLOR:
      local_dispatch<LF_LAMBDA>(kind_, ratio_, v, nullptr, 0, nullptr, nullptr, V_, u00_, s0_, sc, out, nullptr, B, C_, D_, s);
      break;
    case OP_INPAINT:
      inpaint_kernel<LF_LAMBDA><<<blocks(n), 256, 0, s>>>(v, nullptr, 0, nullptr, nullptr, rank_, sc, out, nullptr, B, C_, n2, M_);
      break;
    case OP_WH: {
      float* F = scratch(0, n);
      CUDA_CHECK(cudaMemcpyAsync(F, v, n * 4, cudaMemcpyDeviceToDevice, s));
      fwht(F, B, s);
      wh_spec_kernel<LF_LAMBDA><<<blocks(n), 256, 0, s>>>(F, nullptr, nullptr, perm_, invperm_, ps, out, B, C_, n2, M_);
      fwht(out, B, s);
      break;
    }
    case OP_DEBLUR: {
      float* T = scratch(0, n);
      float* S = scratch(1, n);
      sandwich(Vt_, D_, D_, v, B, V_, D_, D_, T, S, s);
      deblur_coeff_kernel<LF_LAMBDA><<<blocks(n), 256, 0, s>>>(S, nullptr, tabSorig_, ps, S, n2, n);
      sandwich(V_, D_, D_, S, B, Vt_, D_, D_, T, out, s);
      break;
    }
    default:
      throw Error("SRConv defines no Lambda (svd_operators.py:851-931): sigma_y > 0 is unsupported for sr_bicubic, as in the reference");
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::lambda_noise(const float* v, const float* eps, int B, const PlusScalars& ps, float* out, cudaStream_t s) {
  StepScalars sc{};
  sc.plus = ps;
  const int n2 = D_ * D_;
  const long long n = (long long)B * C_ * n2;
  const long long img = (long long)C_ * n2;
  if (kind_ == OP_DENOISE) {
    denoise_kernel<LF_NOISE><<<blocks(n), 256, 0, s>>>(v, nullptr, 0, nullptr, nullptr, sc, out, nullptr, B, img);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  if (kind_ == OP_DEBLUR2D) throw Error("Deblurring2D defines no Lambda_noise (svd_operators.py:1094-1166)");
  if (kind_ == OP_CS) throw Error("CS defines no Lambda_noise (svd_operators.py:101-159)");
  if (kind_ == OP_GENERAL) throw Error("GeneralA defines no Lambda_noise (svd_operators.py:173-208)");
  if (kind_ == OP_SR && sr_generic_) {
    srg_lambda_noise(v, eps, B, ps, out, s);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  switch (kind_) {
    case OP_SR: case OP_COLOR:
      local_dispatch<LF_NOISE>(kind_, ratio_, v, eps, img, nullptr, nullptr, V_, u00_, s0_, sc, out, nullptr, B, C_, D_, s);
      break;
    case OP_INPAINT:
      inpaint_kernel<LF_NOISE><<<blocks(n), 256, 0, s>>>(v, eps, img, nullptr, nullptr, rank_, sc, out, nullptr, B, C_, n2, M_);
      break;
    case OP_WH:
      wh_spec_kernel<LF_NOISE><<<blocks(n), 256, 0, s>>>(v, eps, nullptr, perm_, invperm_, ps, out, B, C_, n2, M_);
      fwht(out, B, s);
      break;
    case OP_DEBLUR: {
      float* T = scratch(0, n);
      float* S = scratch(1, n);
      deblur_coeff_kernel<LF_NOISE><<<blocks(n), 256, 0, s>>>(v, eps, tabSorig_, ps, S, n2, n);
      sandwich(V_, D_, D_, S, B, Vt_, D_, D_, T, out, s);
      break;
    }
    default:
      throw Error("SRConv defines no Lambda_noise (svd_operators.py:851-931)");
  }
  CUDA_CHECK(cudaGetLastError());
}

void Operator::step(const float* xt, const float* et, long long et_stride, const float* noise, const float* y, int B,
                    const StepScalars& sc, float* x0_t, float* xt_next, cudaStream_t s) {
  const int n2 = D_ * D_;
  const long long img = N_;
  const long long n = (long long)B * img;
  if ((kind_ == OP_SR && !sr_generic_) || kind_ == OP_COLOR) {
    local_dispatch<LF_STEP>(kind_, ratio_, xt, et, et_stride, noise, y, V_, u00_, s0_, sc, x0_t, xt_next, B, C_, D_, s);
  } else if (kind_ == OP_INPAINT) {
    inpaint_kernel<LF_STEP><<<blocks(n), 256, 0, s>>>(xt, et, et_stride, noise, y, rank_, sc, x0_t, xt_next, B, C_, n2, M_);
  } else if (kind_ == OP_DENOISE) {
    denoise_kernel<LF_STEP><<<blocks(n), 256, 0, s>>>(xt, et, et_stride, noise, y, sc, x0_t, xt_next, B, img);
  } else {
    // generic path: x0_t, residual r = A^+(A x0_t - y), then the DDNM / DDNM+ update
    float* et3 = scratch(5, n);
    x0_kernel<<<blocks(n), 256, 0, s>>>(xt, et, et_stride, sc, x0_t, et3, B, img);
    float* R = scratch(4, n);
    if (kind_ == OP_WH) {
      float* F = scratch(0, n);
      CUDA_CHECK(cudaMemcpyAsync(F, x0_t, n * 4, cudaMemcpyDeviceToDevice, s));
      fwht(F, B, s);
      wh_spec_kernel<LF_PROJECT><<<blocks(n), 256, 0, s>>>(F, nullptr, y, perm_, invperm_, sc.plus, R, B, C_, n2, M_);
      fwht(R, B, s);
    } else {
      float* Ay = scratch(3, (size_t)B * M_);
      if (kind_ == OP_SR) {          // generic-ratio SuperResolution
        srg_A(x0_t, B, Ay, s);
        sub_kernel<<<blocks((long long)B * M_), 256, 0, s>>>(Ay, y, Ay, (long long)B * M_);
        srg_Apinv(Ay, B, R, s);
      } else {
        deblur_A(x0_t, B, Ay, s);
        sub_kernel<<<blocks((long long)B * M_), 256, 0, s>>>(Ay, y, Ay, (long long)B * M_);
        deblur_Apinv(Ay, B, R, s);
      }
    }
    if (!sc.use_plus) {
      final_ddnm_kernel<<<blocks(n), 256, 0, s>>>(x0_t, R, noise, et3, sc, xt_next, n);
    } else {
      lambda(R, B, sc.plus, R, s);                       // R <- Lambda(R)   (in place is safe: inputs are staged first)
      float* NZ = scratch(3, std::max<size_t>((size_t)n, (size_t)B * M_));
      lambda_noise(noise, et3, B, sc.plus, NZ, s);
      final_plus_kernel<<<blocks(n), 256, 0, s>>>(x0_t, R, NZ, sc, xt_next, n);
    }
  }
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddnm

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
using namespace ddnm;
extern "C" {

int ddnm_operator_create(const ddnm_operator_desc* d, void** handle) {
  DDNM_API_BEGIN
  DDNM_CHECK(d && handle, "null argument");
  *handle = new Operator(d->kind, d->channels, d->img_dim, d->ratio, d->v_small, d->u_small, d->singulars, d->singulars_orig,
                         d->perm, d->mask, d->v_small2, d->u_small2);
  DDNM_API_END
}
long long ddnm_operator_y_dim(void* h) { return h ? static_cast<Operator*>(h)->y_dim() : -1; }
int ddnm_operator_A(void* h, const float* x, int B, float* y, void* stream) {
  DDNM_API_BEGIN
  static_cast<Operator*>(h)->A(x, B, y, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_operator_A_pinv(void* h, const float* y, int B, float* x, void* stream) {
  DDNM_API_BEGIN
  static_cast<Operator*>(h)->A_pinv(y, B, x, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_operator_project(void* h, const float* x0, const float* y, int B, float* out, void* stream) {
  DDNM_API_BEGIN
  static_cast<Operator*>(h)->project(x0, y, B, out, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_operator_lambda(void* h, const float* v, int B, float a, float sigma_y, float sigma_t, float eta, float* out, void* stream) {
  DDNM_API_BEGIN
  static_cast<Operator*>(h)->lambda(v, B, Operator::make_plus(a, sigma_y, sigma_t, eta), out, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_operator_lambda_noise(void* h, const float* v, const float* eps, int B, float a, float sigma_y, float sigma_t, float eta,
                               float* out, void* stream) {
  DDNM_API_BEGIN
  static_cast<Operator*>(h)->lambda_noise(v, eps, B, Operator::make_plus(a, sigma_y, sigma_t, eta), out, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_operator_destroy(void* h) {
  DDNM_API_BEGIN
  delete static_cast<Operator*>(h);
  DDNM_API_END
}

}  // extern "C"
