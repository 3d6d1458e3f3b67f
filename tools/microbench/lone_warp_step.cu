// How fast can ONE warp run the 8-state add-compare-select recursion (no memory traffic)? cycles per trellis step.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t paddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ void beta_step(uint32_t (&b)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = padd(x, y);
  const uint32_t t2 = padd(b[1], x), t3 = padd(b[1], y), t4 = padd(b[2], y), t5 = padd(b[2], x);
  const uint32_t n1 = paddmax(b[0], xy, b[4]), n0 = paddmax(b[4], xy, b[0]), n6 = paddmax(b[3], xy, b[7]), n7 = paddmax(b[7], xy, b[3]);
  const uint32_t n2 = paddmax(b[5], y, t2), n3 = paddmax(b[5], x, t3), n4 = paddmax(b[6], x, t4), n5 = paddmax(b[6], y, t5);
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}
__device__ __forceinline__ void normalise(uint32_t (&s)[8])
{
  uint32_t neg = __vsub2(0u, s[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = padd(s[i], neg);
  s[0] = 0u;
}
// variant: 0 = recursion only, 1 = + normalise every 4 steps, 2 = + x,y from shared memory (loaded at point of use),
// 3 = + shared-memory loads software-pipelined one group of 4 steps ahead
template <int VAR>
__global__ void k2(uint32_t* out, long long* cyc, int iters)
{
  __shared__ uint32_t sx[64][32], sy[64][32];
  const int lane = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < 64; r += blockDim.x >> 5) { sx[r][lane] = r * 3 + lane; sy[r][lane] = r * 5 - lane; }
  __syncthreads();
  uint32_t b[8];
  for (int i = 0; i < 8; i++) b[i] = threadIdx.x * 7 + i * 3;
  uint32_t px[4], py[4];
#pragma unroll
  for (int u = 0; u < 4; u++) { px[u] = sx[u][lane]; py[u] = sy[u][lane]; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int r0 = (it * 4) & 63;
    uint32_t x[4], y[4];
    if (VAR == 3) {
#pragma unroll
      for (int u = 0; u < 4; u++) { x[u] = px[u]; y[u] = py[u]; }
      const int r1 = (r0 + 4) & 63;
#pragma unroll
      for (int u = 0; u < 4; u++) { px[u] = sx[r1 + u][lane]; py[u] = sy[r1 + u][lane]; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t xx, yy;
      if (VAR <= 1) { xx = 3 + u; yy = 5 ^ (uint32_t)it; }
      else if (VAR == 2) { xx = sx[r0 + u][lane]; yy = sy[r0 + u][lane]; }
      else { xx = x[u]; yy = y[u]; }
      beta_step(b, xx, yy);
    }
    if (VAR >= 1) normalise(b);
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ px[0] ^ py[1];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VAR> void run2(const char* what)
{
  uint32_t* out; long long* cyc; cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8 * 148);
  k2<VAR><<<148, 128>>>(out, cyc, 100);
  k2<VAR><<<148, 128>>>(out, cyc, 20000);
  long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  double m = 0; for (int i = 0; i < 148; i++) m += h[i]; m /= 148;
  printf("lone warp per SM sub-partition, %-58s : %.1f cycles per step\n", what, m / 80000.0);
  cudaFree(out); cudaFree(cyc);
}
template <int NCH, int WARPS>
__global__ void k(uint32_t* out, long long* cyc, int iters, uint32_t x0, uint32_t y0)
{
  uint32_t b[NCH][8];
  for (int c = 0; c < NCH; c++)
    for (int i = 0; i < 8; i++) b[c][i] = threadIdx.x * 7 + i * 3 + c;
  uint32_t x = x0, y = y0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < NCH; c++) beta_step(b[c], x + u, y ^ (uint32_t)it);
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int c = 0; c < NCH; c++)
    for (int i = 0; i < 8; i++) acc ^= b[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NCH, int WARPS> void run()
{
  uint32_t* out; long long* cyc; cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8 * 148);
  k<NCH, WARPS><<<148, 32 * WARPS>>>(out, cyc, 100, 3, 5);
  k<NCH, WARPS><<<148, 32 * WARPS>>>(out, cyc, 20000, 3, 5);
  long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  double m = 0; for (int i = 0; i < 148; i++) m += h[i]; m /= 148;
  printf("chains/thread=%d warps/SM=%2d : %.1f cycles per (step of all chains of a warp) -> %.1f cycles per chain-step, SM-wide %.2f chain-steps/cycle\n", NCH, WARPS,
         m / 80000.0, m / 80000.0 / NCH, (double)NCH * WARPS * 80000.0 / m);
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  run2<0>("recursion only"); run2<1>("+ normalise every 4 steps"); run2<2>("+ x,y from shared memory at point of use"); run2<3>("+ shared loads one 4-step group ahead");
  run<1, 4>(); run<2, 4>(); run<1, 8>(); run<1, 16>(); run<2, 8>(); run<1, 32>();
  return 0;
}
