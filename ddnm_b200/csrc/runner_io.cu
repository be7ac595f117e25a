// The runner's I/O step either side of the sampling loop, fused on the device:
//   data_transform / inverse_data_transform      datasets/__init__.py:201-227
//   per-image PSNR against the ground truth      guided_diffusion/diffusion.py:599-602
//   the uint8 HWC quantisation of tvu.save_image guided_diffusion/diffusion.py:596-598 (torchvision: mul(255).add_(0.5).clamp_(0,255).to(uint8))
// One pass over the restored batch yields the [0,1] images, the bytes a PNG encoder needs (4x less D2H than fp32) and the PSNRs;
// the reference does this with ~10 ATen launches and one blocking .to(device) per image.
#include "api_util.cuh"
#include "common.cuh"
#include "../../include/ddnm_b200.h"

namespace ddnm {

__device__ __forceinline__ float inv_transform(float v, int rescaled, int logit) {
  if (logit) v = 1.0f / (1.0f + expf(-v));                 // torch.sigmoid
  else if (rescaled) v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);   // (X + 1.0) / 2.0
  return fminf(fmaxf(v, 0.0f), 1.0f);                      // torch.clamp(X, 0.0, 1.0)
}

__global__ void data_transform_kernel(const float* __restrict__ x, const float* __restrict__ un, const float* __restrict__ gn,
                                      int rescaled, int logit, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (un) v = __fadd_rn(__fmul_rn(__fdiv_rn(v, 256.0f), 255.0f), __fdiv_rn(un[i], 256.0f));   // X / 256.0 * 255.0 + rand / 256.0
  if (gn) v = __fadd_rn(v, __fmul_rn(gn[i], 0.01f));                                           // X + randn * 0.01
  if (rescaled) {
    v = __fsub_rn(__fmul_rn(2.0f, v), 1.0f);                                                   // 2 * X - 1.0
  } else if (logit) {
    const float lam = 1e-6f;
    v = __fadd_rn(lam, __fmul_rn(1.0f - 2.0f * lam, v));                                       // lam + (1 - 2 * lam) * image
    v = __fsub_rn(logf(v), log1pf(-v));                                                        // log(image) - log1p(-image)
  }
  out[i] = v;
}

__global__ void inverse_transform_kernel(const float* __restrict__ x, int rescaled, int logit, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = inv_transform(x[i], rescaled, logit);
}

// grid (chunks, B); thread = one pixel, loops the C planes (coalesced NCHW reads), writes C adjacent bytes of the HWC image.
// partial[b * chunks + chunk] = sum over the chunk of (x01 - orig01)^2 in double; reduced in fixed order by psnr_kernel.
constexpr int FIN_THREADS = 256;
__global__ void __launch_bounds__(FIN_THREADS) finish_kernel(const float* __restrict__ x, const float* __restrict__ orig, int C, int HW,
                                                             int rescaled, int logit, float* __restrict__ out01,
                                                             unsigned char* __restrict__ out_u8, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const long long base = (long long)b * C * HW;
  double acc = 0.0;
  for (int p = blockIdx.x * FIN_THREADS + threadIdx.x; p < HW; p += gridDim.x * FIN_THREADS) {
    for (int c = 0; c < C; ++c) {
      const long long i = base + (long long)c * HW + p;
      const float v = inv_transform(x[i], rescaled, logit);
      if (out01) out01[i] = v;
      if (out_u8) {
        const float q = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f), 0.0f), 255.0f);
        out_u8[((long long)b * HW + p) * C + c] = (unsigned char)q;     // float -> uint8 truncates, as Tensor.to(torch.uint8)
      }
      if (orig) {
        const float d = __fsub_rn(v, inv_transform(orig[i], rescaled, logit));
        acc += (double)__fmul_rn(d, d);
      }
    }
  }
  if (!partial) return;
  __shared__ double red[FIN_THREADS / 32];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < FIN_THREADS / 32; ++w) t += red[w];
    partial[(long long)b * gridDim.x + blockIdx.x] = t;
  }
}

// mse = mean((x - orig) ** 2); psnr = 10 * log10(1 / mse)   (diffusion.py:600-601)
__global__ void psnr_kernel(const double* __restrict__ partial, int chunks, long long elems, float* __restrict__ psnr, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double t = 0.0;
  for (int k = 0; k < chunks; ++k) t += partial[(long long)b * chunks + k];
  const float mse = (float)(t / (double)elems);
  psnr[b] = __fmul_rn(10.0f, log10f(__fdiv_rn(1.0f, mse)));
}

}  // namespace ddnm

using namespace ddnm;
extern "C" {

int ddnm_data_transform(const float* x, long long n, const float* uniform_noise, const float* gauss_noise, int rescaled, int logit,
                        float* out, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(x && out && n >= 0, "null argument");
  if (n == 0) return 0;
  data_transform_kernel<<<(unsigned)cdivll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, uniform_noise, gauss_noise, rescaled, logit, out, n);
  CUDA_CHECK(cudaGetLastError());
  DDNM_API_END
}

int ddnm_inverse_data_transform(const float* x, long long n, int rescaled, int logit, float* out, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(x && out && n >= 0, "null argument");
  if (n == 0) return 0;
  inverse_transform_kernel<<<(unsigned)cdivll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, rescaled, logit, out, n);
  CUDA_CHECK(cudaGetLastError());
  DDNM_API_END
}

int ddnm_finish_images(const float* x, const float* orig, int B, int C, int H, int W, int rescaled, int logit, float* out01,
                       unsigned char* out_u8_hwc, float* psnr, void* stream) {
  DDNM_API_BEGIN
  DDNM_CHECK(x && B >= 0 && C >= 1 && H >= 1 && W >= 1, "bad arguments");
  DDNM_CHECK((orig == nullptr) == (psnr == nullptr), "orig and psnr go together");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int chunks = std::max(1, std::min(cdiv(HW, FIN_THREADS), 64));
  double* partial = nullptr;
  if (orig) CUDA_CHECK(cudaMallocAsync(&partial, sizeof(double) * (size_t)B * chunks, st));
  finish_kernel<<<dim3(chunks, B), FIN_THREADS, 0, st>>>(x, orig, C, HW, rescaled, logit, out01, out_u8_hwc, partial);
  CUDA_CHECK(cudaGetLastError());
  if (orig) {
    psnr_kernel<<<cdiv(B, 128), 128, 0, st>>>(partial, chunks, (long long)C * HW, psnr, B);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaFreeAsync(partial, st));
  }
  DDNM_API_END
}

}  // extern "C"
