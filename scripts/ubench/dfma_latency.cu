// micro-benchmark: dependent DFMA chain latency, DFMA issue rate of one warp, LDS.64 round trip (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int n) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);  // dependent chain
  long long t1 = clock64();
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  for (int i = 0; i < n; ++i) {  // 8 independent chains
    x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
    x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
  }
  long long t2 = clock64();
  int idx = threadIdx.x & 1023;
  double s = 0;
  for (int i = 0; i < n; ++i) { s += sm[idx]; idx = (idx + (int)s) & 1023; }  // dependent LDS + DADD + convert
  long long t3 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + s;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
  double *out; long long *cyc, h[3];
  cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8); cudaMalloc(&cyc, 24);
  const int n = 4096;
  for (int threads : {32, 128, 256}) {
    k<<<1, threads>>>(out, cyc, n); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, n); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("threads %3d: dependent DFMA %.1f cycles each; 8 independent chains %.1f cycles per DFMA; LDS+DADD+cvt loop %.1f cycles per trip\n",
           threads, (double)h[0] / n, (double)h[1] / (8.0 * n), (double)h[2] / n);
  }
  return 0;
}
