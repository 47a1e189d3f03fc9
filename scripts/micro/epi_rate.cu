// Microbenchmark: the P-tile epilogue math of the backward kernel (32-column chunks: FFMA, EX2, FADD, FMUL, F2FP pack)
// run by 8 warps per SM with nothing else on the SM.  Prints cycles per 128 x 256 tile-equivalent (128 elements per
// thread).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_rate epi_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int VARIANT>
__global__ void __launch_bounds__(256, 1) k(uint4* out, long long* cyc, int tiles, float kk, float nshift, float a_i) {
  __shared__ float cvs[256];
  cvs[threadIdx.x] = threadIdx.x * 1e-4f;
  __syncthreads();
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(0.01f * i + threadIdx.x * 1e-3f);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t packed[16];
      const float4* cv4 = reinterpret_cast<const float4*>(cvs + c * 32);
      if (VARIANT == 0) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 cc = cv4[q >> 2];
          float e0 = ex2(fmaf(__uint_as_float(v[q + 0]), kk, nshift)) * (a_i + cc.x);
          float e1 = ex2(fmaf(__uint_as_float(v[q + 1]), kk, nshift)) * (a_i + cc.y);
          float e2 = ex2(fmaf(__uint_as_float(v[q + 2]), kk, nshift)) * (a_i + cc.z);
          float e3 = ex2(fmaf(__uint_as_float(v[q + 3]), kk, nshift)) * (a_i + cc.w);
          __half2 h0 = __floats2half2_rn(e0, e1), h1 = __floats2half2_rn(e2, e3);
          packed[(q >> 1) + 0] = *reinterpret_cast<uint32_t*>(&h0);
          packed[(q >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
      } else {   // no pack: sum instead
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 cc = cv4[q >> 2];
          float e0 = ex2(fmaf(__uint_as_float(v[q + 0]), kk, nshift)) * (a_i + cc.x);
          float e1 = ex2(fmaf(__uint_as_float(v[q + 1]), kk, nshift)) * (a_i + cc.y);
          float e2 = ex2(fmaf(__uint_as_float(v[q + 2]), kk, nshift)) * (a_i + cc.z);
          float e3 = ex2(fmaf(__uint_as_float(v[q + 3]), kk, nshift)) * (a_i + cc.w);
          packed[(q >> 1) + 0] = __float_as_uint(e0 + e1);
          packed[(q >> 1) + 1] = __float_as_uint(e2 + e3);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= packed[i];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += (acc & 1);      // data dependence on the loop so nothing is hoisted
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345678u) out[threadIdx.x] = make_uint4(acc, v[0], v[1], v[2]);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  uint4* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 148 * 8);
  for (int var = 0; var < 2; ++var) {
    const int tiles = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      if (var == 0) k<0><<<148, 256>>>(out, cyc, tiles, 48.f, -10.f, 0.3f);
      else k<1><<<148, 256>>>(out, cyc, tiles, 48.f, -10.f, 0.3f);
    }
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("variant %d (%s): %.0f cycles per tile (8 warps x 128 elements per thread); MUFU bound 2048\n", var,
           var == 0 ? "with F2FP pack" : "no pack", c / tiles);
  }
  return 0;
}
