// Microbenchmark: per-SM throughput of MUFU.EX2, F2FP pack, FFMA and a degree-3 FMA-pipe exp2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  unsigned pk[4] = {0, 0, 0, 0};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 2) { __half2 h = __floats2half2_rn(a[i], a[(i + 1) & 7]); pk[i & 3] ^= *reinterpret_cast<unsigned*>(&h); a[i] += 1.0f; }
      if (MODE == 3) {   // FMA-pipe exp2 of the fractional part (degree-3) + exponent insertion via integer add
        float x = a[i];
        float fl = floorf(x);
        float f = x - fl;
        float p = fmaf(f, 0.0555041f, 0.2402265f);
        p = fmaf(p, f, 0.6931472f);
        p = fmaf(p, f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + ((int)fl << 23)) * 1e-3f;
      }
      if (MODE == 4) a[i] = ex2(fmaf(a[i], 1.0001f, 0.5f)) * (a[(i + 1) & 7] + 0.25f);   // the P-tile epilogue's mix
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + pk[0] + pk[1] + pk[2] + pk[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double ops = (double)iters * 8 * threads;
  printf("%-28s threads/SM %4d: %.2f lane-ops/clk/SM\n", name, threads, ops / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("MUFU.EX2", th);
    run<1>("FFMA", th);
    run<2>("F2FP pack (+FADD,LOP)", th);
    run<3>("poly exp2 (FMA pipe)", th);
    run<4>("FFMA+EX2+FADD+FMUL", th);
  }
  return 0;
}
