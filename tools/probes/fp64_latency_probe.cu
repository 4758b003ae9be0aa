// Probe: FP64 pipe on B200 — dependent-issue latency of DFMA / DADD / DMUL / MUFU.RCP64H and the pipe throughput as a
// function of (warps per SM sub-partition) x (independent chains per thread).  Answers "how much ILP does a 4-warp-per-
// scheduler kernel need to fill the FP64 pipe".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_latency_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x + q;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_add(double* out, long long* cyc, int iters, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = x + b;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rcp(double* out, long long* cyc, int iters) {
  double x = threadIdx.x + 1.5;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_lds(double* out, long long* cyc, int iters) {
  __shared__ int idx[128];
  idx[threadIdx.x] = (threadIdx.x * 0 + threadIdx.x) & 127;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = idx[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps_per_smsp, double* out, long long* cyc) {
  const int iters = 2000;
  const int threads = 128 * warps_per_smsp;      // 4 SMSPs x warps x 32
  chains<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chains<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double nf = 148.0 * threads * (double)iters * 16 * ILP;
  const double per_warp_instr = (double)iters * 16 * ILP;
  printf("warps/SMSP %d ILP %d : %7.2f TFLOP/s, %6.2f cycles per DFMA per warp, pipe cycles per warp-instr at SMSP = %.2f\n", warps_per_smsp, ILP,
         2 * nf / (ms * 1e-3) / 1e12, (double)h / per_warp_instr, (double)h / (per_warp_instr * warps_per_smsp));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  long long h;
  chains<1><<<1, 32>>>(out, cyc, 1000, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", h / 16000.0);
  lat_add<<<1, 32>>>(out, cyc, 1000, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DADD dependent latency: %.2f cycles\n", h / 16000.0);
  lat_rcp<<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H dependent latency: %.2f cycles\n", h / 16000.0);
  lat_lds<<<1, 32>>>(out, cyc, 1000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent latency: %.2f cycles\n", h / 16000.0);
  for (int w : {1, 2, 4, 8}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
