// Throughput of MUFU.TANH / MUFU.EX2 / MUFU.RCP / FFMA / FFMA2 per SM on this GPU (ops per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mufu tools/micro/mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void kern(float* out, int iters, long long* cycles) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
    if (OP == 4) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        uint64_t p;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(v[i]), "f"(v[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(p));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int threads, int ops_per_iter) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  kern<OP><<<148, threads>>>(out, iters, cyc);
  kern<OP><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-10s threads/SM %4d: %.1f ops/clk/SM\n", name, threads, (double)threads * ops_per_iter * iters / h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 512, 1024}) {
    run<0>("MUFU.TANH", th, 8);
    run<1>("MUFU.EX2", th, 8);
    run<2>("MUFU.RCP", th, 8);
    run<3>("FFMA", th, 8);
    run<4>("FFMA2(x2)", th, 16);
  }
  return 0;
}
