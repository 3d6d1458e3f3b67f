// Dependent-issue latency of the packed s16x2 instructions on sm_100a: one warp, one dependent chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int16x2_latency int16x2_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
enum { L_VADD2, L_VIADDMAX2, L_VMAX2, L_ADD_THEN_ADDMAX, L_IADD, L_LOP3, L_ADDMAX_THEN_ADD, L_COUNT };
static const char* names[L_COUNT] = {"VIADD.16x2 -> VIADD.16x2", "VIADDMNMX.S16x2 -> VIADDMNMX.S16x2", "VIMNMX.S16x2 -> VIMNMX.S16x2",
                                     "VIADD.16x2 -> VIADDMNMX.S16x2 -> (pair)", "IADD3 -> IADD3", "LOP3 -> LOP3", "two chains interleaved (VIADD, VIADDMNMX)x2"};
template <int OP>
__global__ void lat(unsigned* out, long long* cyc, int iters, unsigned b, unsigned c)
{
  unsigned a = threadIdx.x * 3u + 1u, a2 = threadIdx.x * 5u + 7u;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 64; u++) {
      if (OP == L_VADD2) a = __vadd2(a, b);
      if (OP == L_VIADDMAX2) a = __viaddmax_s16x2(a, b, c);
      if (OP == L_VMAX2) a = __vmaxs2(a, (u & 1) ? b : c);
      if (OP == L_ADD_THEN_ADDMAX) { unsigned t = __vadd2(a, b); a = __viaddmax_s16x2(a, c, t); }
      if (OP == L_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
      if (OP == L_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
      if (OP == L_ADDMAX_THEN_ADD) {
        unsigned t = __vadd2(a, b), t2 = __vadd2(a2, c);
        a = __viaddmax_s16x2(a, c, t); a2 = __viaddmax_s16x2(a2, b, t2);
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a ^ a2;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run()
{
  unsigned* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  lat<OP><<<1, 32>>>(out, cyc, 10, 0x00010002u, 0x80008000u);
  lat<OP><<<1, 32>>>(out, cyc, 1000, 0x00010002u, 0x80008000u);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-48s %.2f cycles per loop body unit (64 units x 1000)\n", names[OP], (double)h / 64000.0);
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  run<L_VADD2>(); run<L_VIADDMAX2>(); run<L_VMAX2>(); run<L_ADD_THEN_ADDMAX>(); run<L_IADD>(); run<L_LOP3>(); run<L_ADDMAX_THEN_ADD>();
  return 0;
}
