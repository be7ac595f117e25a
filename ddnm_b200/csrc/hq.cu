// hq_demo's arbitrary-size DDNM ("mask-shift" restoration): the per-step arithmetic of
//   hq_demo/guided_diffusion/gaussian_diffusion.py:318-390 (p_mean_variance's "DDNM core": x0_t from eps, clipping, Eq. 17
//   x0_t_hat = lambda_t*Apy + x0_t - lambda_t*Ap(A(x0_t)), the mask-shift overwrite from the canvas, the posterior mean with
//   variance = gamma_t), :431-493 (p_sample, incl. the classifier's condition_mean :414-430) and :208-217 (_undo),
// plus the canvas preparation Ap(A_temp(gt)) of :651-655 for an arbitrary H x W.  The window / time loops stay on the host
// (ddnm_b200/hq.py), as in the reference; every tensor operation of a step runs here.
#include <cmath>

#include "../../include/ddnm_b200.h"
#include "api_util.cuh"
#include "common.cuh"

namespace ddnm {
// simplified.cu: A / Ap of the image-space operators on (B, 3, D, D) images
void simplified_A(const ddnm_simple_deg* d, const float* x, int B, float* y, cudaStream_t st);
void simplified_Ap(const ddnm_simple_deg* d, const float* y, int B, float* x, cudaStream_t st);

// x0_t = clamp(sqrt_recip_alphas_cumprod*x - sqrt_recipm1_alphas_cumprod*eps, -1, 1)       (:404-411, :296-300)
__global__ void hq_x0_kernel(const float* __restrict__ x, const float* __restrict__ mo, long long mo_stride, float c_recip, float c_recipm1,
                             int clip, float* __restrict__ x0, long long img, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long b = i / img, r = i - b * img;
  float v = __fsub_rn(__fmul_rn(c_recip, x[i]), __fmul_rn(c_recipm1, mo[b * mo_stride + r]));
  if (clip) v = fminf(fmaxf(v, -1.0f), 1.0f);
  x0[i] = v;
}

struct HqRect { int dy, dx, h, w, sy, sx; };   // x0_hat[:, :, dy:dy+h, dx:dx+w] = canvas[:, :, sy:sy+h, sx:sx+w]

// x0_hat = lambda*Apy + x0_t - lambda*ApA; mask-shift overwrite; mean = coef1*x0_hat + coef2*x (+ gamma*grad);
// x_next = mean + nonzero*sqrt(gamma)*noise
__global__ void hq_combine_kernel(const float* __restrict__ x, const float* __restrict__ x0t, const float* __restrict__ apa,
                                  const float* __restrict__ apy, const float* __restrict__ canvas, int cH, int cW, HqRect r0, HqRect r1,
                                  const float* __restrict__ grad, const float* __restrict__ z, ddnm_hq_scalars s,
                                  float* __restrict__ x0hat, float* __restrict__ xn, int C, int D, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int px = (int)(i % D), py = (int)((i / D) % D);
  const long long bc = i / ((long long)D * D);
  float v = __fsub_rn(__fadd_rn(__fmul_rn(s.lambda_t, apy[i]), x0t[i]), __fmul_rn(s.lambda_t, apa[i]));
  // the second rectangle is applied after the first (reference order :363-384), so it wins where they overlap
  if (r0.h > 0 && py >= r0.dy && py < r0.dy + r0.h && px >= r0.dx && px < r0.dx + r0.w)
    v = canvas[(bc * cH + (r0.sy + py - r0.dy)) * cW + (r0.sx + px - r0.dx)];
  if (r1.h > 0 && py >= r1.dy && py < r1.dy + r1.h && px >= r1.dx && px < r1.dx + r1.w)
    v = canvas[(bc * cH + (r1.sy + py - r1.dy)) * cW + (r1.sx + px - r1.dx)];
  x0hat[i] = v;
  float mean = __fadd_rn(__fmul_rn(s.coef1, v), __fmul_rn(s.coef2, x[i]));
  if (grad) mean = __fadd_rn(mean, __fmul_rn(s.gamma_t, grad[i]));
  xn[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(s.nonzero, sqrtf(s.gamma_t)), z[i]));
}

// x = sqrt(1 - beta)*x + sqrt(beta)*noise            (:211-217)
__global__ void hq_undo_kernel(float* __restrict__ x, const float* __restrict__ z, float a, float b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, z[i]));
}

// canvas preparation: Apy_temp = Ap(A_temp(gt)) for gt (B, 3, H, W): block means (optionally of the gray image) broadcast back
__global__ void hq_canvas_kernel(const float* __restrict__ gt, float* __restrict__ out, int B, int H, int W, int scale, int use_gray) {
  const int yd = H / scale, xd = W / scale;
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= (long long)B * yd * xd) return;
  const int bx = (int)(blk % xd), by = (int)((blk / xd) % yd), b = (int)(blk / ((long long)xd * yd));
  const float cf = (float)(1.0 / 3.0);
  const float basef = (float)((1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0) + (1.0 / 3.0) * (1.0 / 3.0));
  float acc[3] = {0.f, 0.f, 0.f};
  const long long HW = (long long)H * W;
  for (int k = 0; k < scale * scale; ++k) {
    const long long o = (long long)(by * scale + k / scale) * W + (bx * scale + k % scale);
    float t0 = gt[((long long)b * 3 + 0) * HW + o], t1 = gt[((long long)b * 3 + 1) * HW + o], t2 = gt[((long long)b * 3 + 2) * HW + o];
    if (use_gray) t0 = t1 = t2 = __fadd_rn(__fadd_rn(__fmul_rn(t0, cf), __fmul_rn(t1, cf)), __fmul_rn(t2, cf));
    acc[0] = __fadd_rn(acc[0], t0); acc[1] = __fadd_rn(acc[1], t1); acc[2] = __fadd_rn(acc[2], t2);
  }
  for (int c = 0; c < 3; ++c) {
    float a = scale > 1 ? __fdiv_rn(acc[c], (float)(scale * scale)) : acc[c];
    if (use_gray) a = __fdiv_rn(__fmul_rn(scale > 1 ? __fdiv_rn(acc[0], (float)(scale * scale)) : acc[0], cf), basef);
    for (int k = 0; k < scale * scale; ++k)
      out[((long long)b * 3 + c) * HW + (long long)(by * scale + k / scale) * W + (bx * scale + k % scale)] = a;
  }
}

}  // namespace ddnm

using namespace ddnm;
extern "C" {
int ddnm_hq_canvas(const float* gt, int B, int H, int W, int scale, int use_gray, float* apy_canvas, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(gt && apy_canvas && B >= 1 && scale >= 1 && H % scale == 0 && W % scale == 0, "bad canvas geometry");
  const long long blocks = (long long)B * (H / scale) * (W / scale);
  hq_canvas_kernel<<<(unsigned)cdivll(blocks, 128), 128, 0, (cudaStream_t)stream>>>(gt, apy_canvas, B, H, W, scale, use_gray);
  CUDA_CHECK(cudaGetLastError());
  DDNM_API_END
}

int ddnm_hq_step(const ddnm_simple_deg* deg, const float* x, const float* model_out, int out_ch, const float* apy, const float* canvas,
                 int canvas_h, int canvas_w, const int* rects, const float* grad, const float* noise, const ddnm_hq_scalars* sc, int B,
                 float* x0_hat, float* x_next, float* scratch, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(deg && x && model_out && apy && canvas && rects && noise && sc && x0_hat && x_next && scratch, "null argument");
  DDNM_CHECK(deg->channels == 3 && (out_ch == 3 || out_ch == 6), "hq step: 3-channel images, 3 or 6 model outputs");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = deg->img_dim;
  const long long img = 3LL * D * D, n = (long long)B * img;
  float* x0t = scratch;          // [n]
  float* apa = scratch + n;      // [n]
  float* yb = scratch + 2 * n;   // [<= n]
  hq_x0_kernel<<<(unsigned)cdivll(n, 256), 256, 0, st>>>(x, model_out, (long long)out_ch * D * D, sc->c_recip, sc->c_recipm1, sc->clip, x0t, img, n);
  simplified_A(deg, x0t, B, yb, st);
  simplified_Ap(deg, yb, B, apa, st);
  HqRect r0{rects[0], rects[1], rects[2], rects[3], rects[4], rects[5]}, r1{rects[6], rects[7], rects[8], rects[9], rects[10], rects[11]};
  for (const HqRect& r : {r0, r1})
    if (r.h > 0) DDNM_CHECK(r.w > 0 && r.dy >= 0 && r.dx >= 0 && r.dy + r.h <= D && r.dx + r.w <= D && r.sy >= 0 && r.sx >= 0 &&
                                r.sy + r.h <= canvas_h && r.sx + r.w <= canvas_w, "mask-shift rectangle out of range");
  hq_combine_kernel<<<(unsigned)cdivll(n, 256), 256, 0, st>>>(x, x0t, apa, apy, canvas, canvas_h, canvas_w, r0, r1, grad, noise, *sc, x0_hat,
                                                              x_next, 3, D, n);
  CUDA_CHECK(cudaGetLastError());
  DDNM_API_END
}

int ddnm_hq_undo(float* x, const float* noise, float sqrt_one_minus_beta, float sqrt_beta, long long n, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(x && noise && n > 0, "null argument");
  hq_undo_kernel<<<(unsigned)cdivll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, noise, sqrt_one_minus_beta, sqrt_beta, n);
  CUDA_CHECK(cudaGetLastError());
  DDNM_API_END
}
}
