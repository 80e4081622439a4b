// fp64_bench.cu -- latency / throughput of the FP64 pipe and of shared-memory loads on one SM (diagnostics).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu && ./fp64_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  unsigned idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // dependent DADD chain
#pragma unroll
      for (int u = 0; u < 16; ++u) x0 = __dadd_rn(x0, a);
    } else if (MODE == 1) {  // 8 independent DADD chains
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a);
        x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a);
      }
    } else if (MODE == 2) {  // dependent DFMA chain
#pragma unroll
      for (int u = 0; u < 16; ++u) x0 = fma(x0, b, a);
    } else if (MODE == 3) {  // 8 independent DFMA chains
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
        x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
      }
    } else if (MODE == 4) {  // dependent shared-memory load chain (pointer chase)
#pragma unroll
      for (int u = 0; u < 16; ++u) idx = (unsigned)sm[idx & 4095] + threadIdx.x;
    } else if (MODE == 5) {  // FADD dependent chain for comparison
      float f = (float)x0;
#pragma unroll
      for (int u = 0; u < 16; ++u) f = __fadd_rn(f, (float)a);
      x0 = f;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + idx;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int ops_per_iter) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  const int iters = 2000;
  k<MODE><<<1, threads>>>(out, cyc, iters, 1e-9, 1.0000001);
  cudaDeviceSynchronize();
  k<MODE><<<1, threads>>>(out, cyc, iters, 1e-9, 1.0000001);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters / ops_per_iter;
  printf("%-40s threads %4d: %7.2f cycles per op per thread -> %7.2f warp-instr/clk/SM\n", name, threads, per,
         (threads / 32.0) / per);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int threads : {32, 128, 512, 1024}) {
    run<0>("DADD dependent chain", threads, 16);
    run<1>("DADD 8 independent chains", threads, 32);
    run<2>("DFMA dependent chain", threads, 16);
    run<3>("DFMA 8 independent chains", threads, 32);
    run<4>("LDS.64 dependent chain", threads, 16);
    run<5>("FADD dependent chain", threads, 16);
  }
  return 0;
}
