// Microbenchmark (development aid, not product): global-memory latency seen by one thread on the target GPU --
// cold DRAM, L2 hit, and after prefetch.global.L2 / cp.async.bulk.prefetch.L2 issued well in advance.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int ld_cg(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// mode 0: timed dependent loads.  mode 1: prefetch.global.L2 of every address first, sleep, then timed loads.
// mode 2: cp.async.bulk.prefetch.L2 (one 128-byte chunk per address), sleep, timed loads.
__global__ void lat_kernel(const int* buf, size_t stride, int n, int mode, long long* out) {
  if (threadIdx.x != 0) return;
  if (mode == 1) {
    for (int i = 0; i < n; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + i * stride));
  } else if (mode == 2) {
    for (int i = 0; i < n; i++) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(buf + i * stride), "r"(128) : "memory");
  }
  if (mode == 3) {
    for (int i = 0; i < n; i++) asm volatile("prefetch.global.L1 [%0];" ::"l"(buf + i * stride));
  }
  if (mode != 0 && mode != 4) {
    for (int k = 0; k < 40; k++) __nanosleep(1000);
  }
  long long total = 0;
  size_t idx = 0;
  int sum = 0;
  for (int i = 0; i < n; i++) {
    long long t0 = clock64();
    int v = (mode == 3 || mode == 4) ? __ldg(buf + idx) : ld_cg(buf + idx);  // 3/4: through L1 (ld.global.nc)
    idx += stride + v;  // v == 0: next address depends on the load
    sum += v;
    long long t1 = clock64();
    total += t1 - t0;
  }
  out[0] = total;
  out[1] = sum;
}

int main() {
  const size_t bytes = 256ull << 20;
  int *buf, *flush;
  long long* out;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&flush, bytes);
  cudaMalloc(&out, 16);
  cudaMemset(buf, 0, bytes);
  const int n = 256;
  const size_t stride = (1 << 20) / 4 + 64;  // ~1 MiB apart: different DRAM pages / L2 slices
  const char* names[] = {"plain load", "after prefetch.global.L2", "after cp.async.bulk.prefetch.L2", "ld.nc after prefetch.global.L1", "ld.nc plain"};
  for (int rep = 0; rep < 2; rep++) {
    for (int mode = 0; mode < 5; mode++) {
      cudaMemset(flush, rep + mode, bytes);  // evict buf from L2 (leaves dirty lines of `flush`)
      cudaDeviceSynchronize();
      lat_kernel<<<1, 32>>>(buf, stride, n, mode, out);
      long long h[2];
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("cold L2 (dirty), %-34s: %.0f clk per dependent load\n", names[mode], (double)h[0] / n);
    }
    // L2 hit: same addresses again without flushing
    lat_kernel<<<1, 32>>>(buf, stride, n, 0, out);
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("warm L2,         %-34s: %.0f clk per dependent load\n", names[0], (double)h[0] / n);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
