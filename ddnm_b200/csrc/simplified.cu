// The "simplified" DDNM+ loop of the reference runner (guided_diffusion/diffusion.py:211-415, README quick start):
// image-space operators built from mask / colour-to-gray / average-pool (diffusion.py:244-290, helpers :27-42) and a scalar
// lambda_t / gamma_t update (:355-376).  One fused kernel per step; thread = one scale x scale patch across the 3 channels.
#include <cmath>
#include <memory>

#include "../../include/ddnm_b200.h"
#include "api_util.cuh"
#include "engine.cuh"

namespace ddnm {

struct SimpScalars {
  float sqrt_at, sqrt_1m_at, sqrt_atn, c1, c2, lambda_t, gamma_t;
};
struct SimpDeg {
  int use_mask, use_gray, scale, D;
  const float* mask;
};

enum { SF_A = 0, SF_AP = 1, SF_STEP = 2 };

// A(z) = pool(gray(z * mask)),  Ap(v) = gray2color(upsample(v)) * mask   (whichever stages are enabled)
template <int S, int FN>
__global__ void __launch_bounds__(128) simp_kernel(const float* __restrict__ in0, const float* __restrict__ et, long long et_stride,
                                                   const float* __restrict__ z, const float* __restrict__ y, SimpDeg dg, SimpScalars sc,
                                                   float* __restrict__ out0, float* __restrict__ out1, int B) {
  constexpr int K = S * S;
  const int D = dg.D, yd = D / S;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)B * yd * yd) return;
  const int px = (int)(g % yd), py = (int)((g / yd) % yd), b = (int)(g / ((long long)yd * yd));
  const long long HW = (long long)D * D, img = 3 * HW;
  const float cf = (float)(1.0 / 3.0);
  const float basef = (float)((1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0));
  auto off = [&](int c, int k) { return (long long)c * HW + (long long)(py * S + k / S) * D + (px * S + k % S); };
  float m[K];
#pragma unroll
  for (int k = 0; k < K; ++k) m[k] = dg.use_mask ? __ldg(dg.mask + (long long)(py * S + k / S) * D + (px * S + k % S)) : 1.f;
  const long long yo = (long long)b * 3 * yd * yd + (long long)py * yd + px;   // + c*yd*yd

  if (FN == SF_AP) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = y[yo + (long long)c * yd * yd];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float w = dg.use_gray ? __fdiv_rn(__fmul_rn(v[0], cf), basef) : v[c];
        if (dg.use_mask) w = __fmul_rn(w, m[k]);
        out0[(long long)b * img + off(c, k)] = w;
      }
    return;
  }

  float x0[3][K];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float xv = in0[(long long)b * img + off(c, k)];
      if (FN == SF_STEP) {
        const float e = et[(long long)b * et_stride + off(c, k)];
        x0[c][k] = __fdiv_rn(__fsub_rn(xv, __fmul_rn(e, sc.sqrt_1m_at)), sc.sqrt_at);
        out0[(long long)b * img + off(c, k)] = x0[c][k];
      } else {
        x0[c][k] = xv;
      }
    }
  // A
  float a[3];
  {
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float t0 = x0[0][k], t1 = x0[1][k], t2 = x0[2][k];
      if (dg.use_mask) { t0 = __fmul_rn(t0, m[k]); t1 = __fmul_rn(t1, m[k]); t2 = __fmul_rn(t2, m[k]); }
      if (dg.use_gray) {
        const float gk = __fadd_rn(__fadd_rn(__fmul_rn(t0, cf), __fmul_rn(t1, cf)), __fmul_rn(t2, cf));
        t0 = t1 = t2 = gk;
      }
      acc[0] = __fadd_rn(acc[0], t0); acc[1] = __fadd_rn(acc[1], t1); acc[2] = __fadd_rn(acc[2], t2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) a[c] = K > 1 ? __fdiv_rn(acc[c], (float)K) : acc[c];
  }
  if (FN == SF_A) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out0[yo + (long long)c * yd * yd] = a[c];
    return;
  }
  float r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) r[c] = __fsub_rn(a[c], y[yo + (long long)c * yd * yd]);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float w = dg.use_gray ? __fdiv_rn(__fmul_rn(r[0], cf), basef) : r[c];
      if (dg.use_mask) w = __fmul_rn(w, m[k]);
      const float x0h = __fsub_rn(x0[c][k], __fmul_rn(sc.lambda_t, w));                       // Eq. 17 (:373)
      const float e = et[(long long)b * et_stride + off(c, k)];
      const float zz = z[(long long)b * img + off(c, k)];
      const float nz = __fmul_rn(sc.gamma_t, __fadd_rn(__fmul_rn(sc.c1, zz), __fmul_rn(sc.c2, e)));  // (:381)
      out1[(long long)b * img + off(c, k)] = __fadd_rn(__fmul_rn(sc.sqrt_atn, x0h), nz);
    }
}

// Any other scale (evaluation.sh runs sr_averagepooling with deg_scale 16): one WARP per patch, lanes stride over the S*S pixels,
// the three channel sums meet through shuffles.  Same arithmetic per element as simp_kernel; the pooled sums are re-associated.
template <int FN>
__global__ void __launch_bounds__(128) simp_generic_kernel(const float* __restrict__ in0, const float* __restrict__ et, long long et_stride,
                                                           const float* __restrict__ z, const float* __restrict__ y, SimpDeg dg,
                                                           SimpScalars sc, float* __restrict__ out0, float* __restrict__ out1, int B) {
  const int S = dg.scale, K = S * S;
  const int D = dg.D, yd = D / S;
  const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= (long long)B * yd * yd) return;
  const int px = (int)(g % yd), py = (int)((g / yd) % yd), b = (int)(g / ((long long)yd * yd));
  const long long HW = (long long)D * D, img = 3 * HW;
  const float cf = (float)(1.0 / 3.0);
  const float basef = (float)((1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0));
  auto pix = [&](int k) { return (long long)(py * S + k / S) * D + (px * S + k % S); };
  const long long yo = (long long)b * 3 * yd * yd + (long long)py * yd + px;   // + c*yd*yd
  if (FN == SF_AP) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = y[yo + (long long)c * yd * yd];
    for (int k = lane; k < K; k += 32) {
      const float m = dg.use_mask ? __ldg(dg.mask + pix(k)) : 1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float w = dg.use_gray ? __fdiv_rn(__fmul_rn(v[0], cf), basef) : v[c];
        if (dg.use_mask) w = __fmul_rn(w, m);
        out0[(long long)b * img + (long long)c * HW + pix(k)] = w;
      }
    }
    return;
  }
  float acc[3] = {0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float m = dg.use_mask ? __ldg(dg.mask + pix(k)) : 1.f;
    float t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o = (long long)c * HW + pix(k);
      const float xv = in0[(long long)b * img + o];
      if (FN == SF_STEP) {
        const float e = et[(long long)b * et_stride + o];
        t[c] = __fdiv_rn(__fsub_rn(xv, __fmul_rn(e, sc.sqrt_1m_at)), sc.sqrt_at);
        out0[(long long)b * img + o] = t[c];
      } else {
        t[c] = xv;
      }
      if (dg.use_mask) t[c] = __fmul_rn(t[c], m);
    }
    if (dg.use_gray) {
      const float gk = __fadd_rn(__fadd_rn(__fmul_rn(t[0], cf), __fmul_rn(t[1], cf)), __fmul_rn(t[2], cf));
      t[0] = t[1] = t[2] = gk;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], t[c]);
  }
  float a[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    a[c] = K > 1 ? __fdiv_rn(v, (float)K) : v;
  }
  if (FN == SF_A) {
    if (lane < 3) out0[yo + (long long)lane * yd * yd] = lane == 0 ? a[0] : (lane == 1 ? a[1] : a[2]);
    return;
  }
  float r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) r[c] = __fsub_rn(a[c], y[yo + (long long)c * yd * yd]);
  for (int k = lane; k < K; k += 32) {
    const float m = dg.use_mask ? __ldg(dg.mask + pix(k)) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o = (long long)c * HW + pix(k);
      float w = dg.use_gray ? __fdiv_rn(__fmul_rn(r[0], cf), basef) : r[c];
      if (dg.use_mask) w = __fmul_rn(w, m);
      const float x0 = out0[(long long)b * img + o];                                            // written above by this thread
      const float x0h = __fsub_rn(x0, __fmul_rn(sc.lambda_t, w));                               // Eq. 17 (:373)
      const float e = et[(long long)b * et_stride + o];
      const float zz = z[(long long)b * img + o];
      const float nz = __fmul_rn(sc.gamma_t, __fadd_rn(__fmul_rn(sc.c1, zz), __fmul_rn(sc.c2, e)));  // (:381)
      out1[(long long)b * img + o] = __fadd_rn(__fmul_rn(sc.sqrt_atn, x0h), nz);
    }
  }
}

template <int FN>
static void simp_launch(const SimpDeg& dg, const float* in0, const float* et, long long et_stride, const float* z, const float* y,
                        const SimpScalars& sc, float* out0, float* out1, int B, cudaStream_t st) {
  const int yd = dg.D / dg.scale;
  const long long groups = (long long)B * yd * yd;
  const int grid = (int)cdivll(groups, 128);
  switch (dg.scale) {
    case 1: simp_kernel<1, FN><<<grid, 128, 0, st>>>(in0, et, et_stride, z, y, dg, sc, out0, out1, B); break;
    case 2: simp_kernel<2, FN><<<grid, 128, 0, st>>>(in0, et, et_stride, z, y, dg, sc, out0, out1, B); break;
    case 4: simp_kernel<4, FN><<<grid, 128, 0, st>>>(in0, et, et_stride, z, y, dg, sc, out0, out1, B); break;
    case 8: simp_kernel<8, FN><<<grid, 128, 0, st>>>(in0, et, et_stride, z, y, dg, sc, out0, out1, B); break;
    default:   // any other scale dividing the image size
      simp_generic_kernel<FN><<<(int)cdivll(groups * 32, 128), 128, 0, st>>>(in0, et, et_stride, z, y, dg, sc, out0, out1, B);
  }
  CUDA_CHECK(cudaGetLastError());
}

static SimpDeg make_deg(const ddnm_simple_deg* d) {
  DDNM_CHECK(d != nullptr, "null degradation");
  DDNM_CHECK(d->channels == 3 && d->img_dim > 0 && d->scale >= 1 && d->img_dim % d->scale == 0, "bad simplified degradation");
  DDNM_CHECK(!d->use_mask || d->mask != nullptr, "mask enabled but no mask given");
  SimpDeg g;
  g.use_mask = d->use_mask; g.use_gray = d->use_gray; g.scale = d->scale; g.D = d->img_dim; g.mask = d->mask;
  return g;
}

__global__ void simp_fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void simp_travel_kernel(const float* __restrict__ x0, const float* __restrict__ z, float sa, float s1, float* __restrict__ xn,
                                   long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xn[i] = __fadd_rn(__fmul_rn(sa, x0[i]), __fmul_rn(z[i], s1));
}

// pairs [k0, k1) with the state in the caller's buffers (same contract as sample_range in sampler.cu)
static void sample_simplified_range(UNetEngine* unet, const ddnm_simple_deg* d, const ddnm_schedule* sc, int k0, int k1, float* xt_state,
                                    float* x0t, int* have_x0, const float* y, const float* noise, int B, cudaStream_t st) {
  DDNM_CHECK(unet && sc && xt_state && x0t && have_x0 && y && noise, "null argument");
  DDNM_CHECK(unet->batch() == B, "engine was built for a different batch size");
  DDNM_CHECK(0 <= k0 && k0 <= k1 && k1 <= sc->n_pairs, "pair range outside the schedule");
  SimpDeg dg = make_deg(d);
  const int R = unet->resolution();
  DDNM_CHECK(dg.D == R && unet->in_channels() == 3, "degradation / denoiser image size mismatch");
  const long long n = (long long)B * 3 * R * R;
  const long long et_stride = (long long)unet->out_ch() * R * R;
  float* xt = unet->x_in();
  float* et = unet->out_buf();
  StreamBuf xnb((size_t)n, st);
  float* xn = xnb.p;
  CUDA_CHECK(cudaMemcpyAsync(xt, xt_state, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const float eta = sc->eta, sigma_y = sc->sigma_y;
  const float c_eta = (float)std::sqrt(1.0 - (double)eta * (double)eta);
  for (int k = k0; k < k1; ++k) {
    const int i = sc->t_i[k], j = sc->t_j[k];
    DDNM_CHECK(i >= 0 && i < sc->num_timesteps && j >= -1 && j < sc->num_timesteps, "time index out of range");
    const float at_next = sc->abar[j + 1];
    const float* z = noise + (long long)(k - k0) * n;
    if (j < i) {
      const float at = sc->abar[i + 1];
      simp_fill_kernel<<<cdiv(B, 128), 128, 0, st>>>(unet->t_in(), B, (float)i);
      unet->forward(xt, unet->t_in(), et, st);
      SimpScalars s{};
      s.sqrt_at = std::sqrt(at);
      s.sqrt_1m_at = std::sqrt(1.0f - at);
      s.sqrt_atn = std::sqrt(at_next);
      const float s1n = std::sqrt(1.0f - at_next);
      s.c1 = s1n * eta;
      s.c2 = s1n * c_eta;
      // Eq. 19 with the runner's sigma_t = sqrt(1 - at_next**2)  (diffusion.py:356, :366-371)
      const float sigma_t = std::sqrt(1.0f - at_next * at_next);
      const float asy = at_next * sigma_y;
      if (sigma_t >= asy) {
        s.lambda_t = 1.0f;
        s.gamma_t = std::sqrt(sigma_t * sigma_t - asy * asy);
      } else {
        s.lambda_t = sigma_t / asy;
        s.gamma_t = 0.0f;
      }
      simp_launch<SF_STEP>(dg, xt, et, et_stride, z, y, s, x0t, xn, B, st);
      *have_x0 = 1;
    } else {
      DDNM_CHECK(*have_x0, "schedule starts with a travel-back step");
      simp_travel_kernel<<<(int)cdivll(n, 256), 256, 0, st>>>(x0t, z, std::sqrt(at_next), std::sqrt(1.0f - at_next), xn, n);
      CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaMemcpyAsync(xt, xn, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  CUDA_CHECK(cudaMemcpyAsync(xt_state, xt, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

static void sample_simplified(UNetEngine* unet, const ddnm_simple_deg* d, const ddnm_schedule* sc, const float* x_T, const float* y,
                              const float* noise, int B, float* out_x0, float* out_x0_pred, cudaStream_t st) {
  DDNM_CHECK(unet && sc && x_T && y && noise && out_x0, "null argument");
  DDNM_CHECK(unet->batch() == B, "engine was built for a different batch size");
  const long long n = (long long)B * 3 * unet->resolution() * unet->resolution();
  std::unique_ptr<StreamBuf> own;
  float* x0t = out_x0_pred;
  if (!x0t) {
    own.reset(new StreamBuf((size_t)n, st));
    x0t = own->p;
  }
  if (out_x0 != x_T) CUDA_CHECK(cudaMemcpyAsync(out_x0, x_T, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  int have_x0 = 0;
  sample_simplified_range(unet, d, sc, 0, sc->n_pairs, out_x0, x0t, &have_x0, y, noise, B, st);
}

// used by the hq_demo mask-shift step (hq.cu)
void simplified_A(const ddnm_simple_deg* d, const float* x, int B, float* y, cudaStream_t st) {
  SimpScalars s{};
  simp_launch<SF_A>(make_deg(d), x, nullptr, 0, nullptr, nullptr, s, y, nullptr, B, st);
}
void simplified_Ap(const ddnm_simple_deg* d, const float* y, int B, float* x, cudaStream_t st) {
  SimpScalars s{};
  simp_launch<SF_AP>(make_deg(d), nullptr, nullptr, 0, nullptr, y, s, x, nullptr, B, st);
}

}  // namespace ddnm

using namespace ddnm;
extern "C" {
int ddnm_simplified_A(const ddnm_simple_deg* d, const float* x, int B, float* y, void* stream) {
  DDNM_API_BEGIN
  SimpScalars s{};
  simp_launch<SF_A>(make_deg(d), x, nullptr, 0, nullptr, nullptr, s, y, nullptr, B, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_simplified_Ap(const ddnm_simple_deg* d, const float* y, int B, float* x, void* stream) {
  DDNM_API_BEGIN
  SimpScalars s{};
  simp_launch<SF_AP>(make_deg(d), nullptr, nullptr, 0, nullptr, y, s, x, nullptr, B, (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_sample_simplified_range(void* unet, const ddnm_simple_deg* d, const ddnm_schedule* sched, int k_begin, int k_end, float* xt,
                                 float* x0_pred, int* have_x0, const float* y, const float* noise, int B, void* stream) {
  DDNM_API_BEGIN
  sample_simplified_range(static_cast<UNetEngine*>(unet), d, sched, k_begin, k_end, xt, x0_pred, have_x0, y, noise, B,
                          (cudaStream_t)stream);
  DDNM_API_END
}
int ddnm_sample_simplified(void* unet, const ddnm_simple_deg* d, const ddnm_schedule* sched, const float* x_T, const float* y,
                           const float* noise, int B, float* out_x0, float* out_x0_pred, void* stream) {
  DDNM_API_BEGIN
  sample_simplified(static_cast<UNetEngine*>(unet), d, sched, x_T, y, noise, B, out_x0, out_x0_pred, (cudaStream_t)stream);
  DDNM_API_END
}
}
