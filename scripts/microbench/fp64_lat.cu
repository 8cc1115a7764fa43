// Microbenchmark (development aid, not product): FP64 latency / throughput on the target GPU.
//   dependent DFMA chain latency, independent-chain throughput per SM sub-partition for 1..8 warps, 1.0/x cost.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) v[k] = threadIdx.x + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = fma(v[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void rcp_kernel(double* out, long long* cycles, int iters, double a) {
  double v = 1.5 + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) v = 1.0 / v + a;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int threads, const char* name) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double) * 148);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  const int iters = 4096;
  dfma_kernel<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  dfma_kernel<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  double warp_instr = (double)iters * ILP;                       // per warp
  double per_smsp = warp_instr * (threads / 32) / 4.0;            // warp instructions per SM sub-partition
  printf("%s threads=%4d ILP=%d : %.2f clk per dependent step, %.2f clk per warp-DFMA per SMSP\n", name, threads, ILP, c / iters, c / per_smsp);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<1>(32, "chain ");
  run<2>(32, "chain ");
  run<4>(32, "chain ");
  run<8>(32, "chain ");
  run<1>(128, "1w/smsp");
  run<4>(128, "1w/smsp");
  run<8>(128, "1w/smsp");
  run<16>(128, "1w/smsp");
  run<4>(256, "2w/smsp");
  run<8>(256, "2w/smsp");
  run<4>(512, "4w/smsp");
  run<8>(512, "4w/smsp");
  run<8>(1024, "8w/smsp");
  {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    rcp_kernel<<<1, 32>>>(out, cyc, 1024, 0.25);
    rcp_kernel<<<1, 32>>>(out, cyc, 1024, 0.25);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("1.0/x + a dependent chain: %.1f clk per step\n", (double)h / 1024);
  }
  return 0;
}
