// Microbenchmark: dependent DFMA chains. ILP independent chains per thread, W warps per SM sub-partition.
// Prints cycles per DFMA per warp and FP64 pipe utilisation, to tell how much ILP x TLP the FP64 pipe of a B200 needs.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp) {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  int iters = 20000, threads = 32 * 4 * warps_per_smsp;
  chain<ILP><<<148, threads>>>(d, 100, 0.999, 1e-9, c);
  chain<ILP><<<148, threads>>>(d, iters, 0.999, 1e-9, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0 * ILP);   // cycles per DFMA per warp
  double util = (2.0 * warps_per_smsp) / (per);   // a warp-DFMA occupies the SMSP's FP64 pipe for 2 cycles
  printf("ILP %d warps/SMSP %d: %.2f cycles per DFMA per warp -> chain latency %.1f cycles, FP64 pipe %.0f%%\n", ILP, warps_per_smsp, per, per * ILP, util * 100);
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int w : {1, 2, 4, 8}) run<1>(w);
  for (int w : {1, 2, 4}) run<2>(w);
  for (int w : {1, 2, 4}) run<3>(w);
  for (int w : {1, 4}) run<4>(w);
  return 0;
}
