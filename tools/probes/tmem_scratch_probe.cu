// Probe: tensor memory (TMEM) as a per-thread FP64 scratchpad without any MMA.
// Each CTA of 128 threads allocates 128 columns (64 doubles per thread), every thread parks 64 doubles in its own lane,
// does some arithmetic, reads them back in a different order and checks.  Also times LDTM/STTM round trips.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_scratch_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st(unsigned taddr, double v) {
  unsigned lo = __double2loint(v), hi = __double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tmem_ld(unsigned taddr) {
  unsigned lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 4) probe(double* out, long long* cycles, int iters) {
  __shared__ unsigned base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&base_s)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned base = base_s + ((unsigned)(warp * 32) << 16);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int s = 0; s < 64; ++s) tmem_st(base + 2 * s, (double)gid * 1000.0 + s + it);
    tmem_wait_st();
    for (int s = 63; s >= 0; --s) {
      double v = tmem_ld(base + 2 * s);
      if (v != (double)gid * 1000.0 + s + it) acc += 1e9;      // mismatch marker
      acc += v * 1e-9;
    }
  }
  long long t1 = clock64();
  out[gid] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "n"(128));
}

int main() {
  const int blocks = 148 * 8, threads = 128, iters = 50;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks);
  probe<<<blocks, threads>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  double* h = new double[blocks * threads]; long long* hc = new long long[blocks];
  cudaMemcpy(h, out, sizeof(double) * blocks * threads, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < blocks * threads; ++i) if (h[i] > 1e8) ++bad;
  double mean = 0; for (int b = 0; b < blocks; ++b) mean += hc[b]; mean /= blocks;
  printf("mismatching threads: %d of %d\n", bad, blocks * threads);
  printf("mean cycles per (64 st + 64 ld) round, 4 CTAs/SM resident: %.0f  => %.1f cycles per st+ld pair per warp\n", mean / iters, mean / iters / 64);
  return bad != 0;
}
