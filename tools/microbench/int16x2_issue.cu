// Microbenchmark: issue rate and wrap semantics of the packed s16x2 integer instructions on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int16x2_issue int16x2_issue.cu
// Output: one line per (op, warps/SM): warp-instructions per cycle per SM, lane-ops per second chip-wide.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { OP_VADD2 = 0, OP_VMAX2, OP_VIADDMAX2, OP_VIMAX3_2, OP_IADD3, OP_LOP3, OP_IMAD, OP_MIX_VADD_IMAD, OP_MIX_VADD_VMAX,
       OP_VIADDMAX32, OP_MIX_VIADDMAX2_IMAD, OP_MIX_VADD2_FFMA, OP_PRMT, OP_MIX_VADD2_VIADDMAX2, OP_MIX_VMAX2_IMAD, OP_MIX_VADD2_VIMAX3, OP_MIX_1VADD_2VIADDMAX, OP_MIX_VADD2_VMAX2_INDEP, OP_MIX_VADD2_LOP3, OP_MIX_VIADDMAX2_LOP3, OP_COUNT };
static const char* op_names[OP_COUNT] = {"VIADD.16x2", "VIMNMX.S16x2", "VIADDMNMX.S16x2", "VIMNMX3.S16x2", "IADD3", "LOP3", "IMAD",
  "mix VIADD.16x2+IMAD (1:1)", "mix VIADD.16x2+VIMNMX.S16x2 (1:1)", "VIADDMNMX.S32", "mix VIADDMNMX.S16x2+IMAD (1:1)", "mix VIADD.16x2+FFMA (1:1)", "PRMT", "mix VIADD.16x2+VIADDMNMX.S16x2 (1:1)", "mix VIMNMX.S16x2+IMAD (1:1)", "mix VIADD.16x2+VIMNMX3.S16x2 (1:1)", "mix VIADD.16x2+VIADDMNMX.S16x2 (1:2)", "mix VIADD.16x2+VIMNMX.S16x2 vs const (1:1)", "mix VIADD.16x2+LOP3 (1:1)", "mix VIADDMNMX.S16x2+LOP3 (1:1)"};

template <int OP>
__device__ __forceinline__ void step(unsigned (&a)[8], unsigned b, unsigned c, float (&f)[8], float fb)
{
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (OP == OP_VADD2) a[i] = __vadd2(a[i], b);
    if (OP == OP_VMAX2) a[i] = __vmaxs2(a[i], a[(i + 3) & 7]);
    if (OP == OP_VIADDMAX2) a[i] = __viaddmax_s16x2(a[i], b, c);
    if (OP == OP_VIMAX3_2) a[i] = __vimax3_s16x2(a[i], b, c);
    if (OP == OP_IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (OP == OP_MIX_VADD_IMAD) { if (i & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_MIX_VADD_VMAX) { if (i & 1) a[i] = __vmaxs2(a[i], a[(i + 2) & 7]); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_VIADDMAX32) a[i] = (unsigned)__viaddmax_s32((int)a[i], (int)b, (int)c);
    if (OP == OP_MIX_VIADDMAX2_IMAD) { if (i & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else a[i] = __viaddmax_s16x2(a[i], b, c); }
    if (OP == OP_MIX_VADD2_FFMA) { if (i & 1) f[i] = fmaf(f[i], fb, fb); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_PRMT) a[i] = __byte_perm(a[i], b, 0x5432);
    if (OP == OP_MIX_VADD2_VIADDMAX2) { if (i & 1) a[i] = __viaddmax_s16x2(a[i], b, c); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_MIX_VMAX2_IMAD) { if (i & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else a[i] = __vmaxs2(a[i], a[(i + 2) & 7]); }
    if (OP == OP_MIX_VADD2_VIMAX3) { if (i & 1) a[i] = __vimax3_s16x2(a[i], b, c); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_MIX_1VADD_2VIADDMAX) { if (i % 3 == 0 && i < 6) a[i] = __vadd2(a[i], b); else if (i < 6) a[i] = __viaddmax_s16x2(a[i], b, c); }
    if (OP == OP_MIX_VADD2_VMAX2_INDEP) { if (i & 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_MIX_VADD2_LOP3) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); else a[i] = __vadd2(a[i], b); }
    if (OP == OP_MIX_VIADDMAX2_LOP3) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); else a[i] = __viaddmax_s16x2(a[i], b, c); }
  }
}

template <int OP>
__global__ void bench(unsigned* out, long long* cycles, int iters, unsigned b, unsigned c)
{
  unsigned a[8];
  float    f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 77u + i * 0x10003u; f[i] = (float)i + threadIdx.x; }
  float fb = __uint_as_float(0x3f800001u + b);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) step<OP>(a, b, c, f, fb);
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= a[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int nsm, double clk_ghz)
{
  const int iters = 2000;
  unsigned* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(unsigned) * nsm * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
  int tcs[] = {128, 256, 512, 1024};
  for (int t = 0; t < 4; t++) {
    int threads = tcs[t];
    bench<OP><<<nsm, threads>>>(out, cyc, 10, 0x00010002u, 0x80008000u);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<nsm, threads>>>(out, cyc, iters, 0x00010002u, 0x80008000u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < nsm; i++) mean += h[i]; mean /= nsm;
    double winstr = (double)iters * (OP == OP_MIX_1VADD_2VIADDMAX ? 48 : 64) * (threads / 32);  // warp-instructions per SM
    printf("%-36s warps/SM=%2d  cycles=%9.0f  warp-instr/cycle/SM=%6.3f  (%5.1f lanes/clk/SM)  chip %.2f T lane-instr/s (event %.3f ms)\n",
           op_names[OP], threads / 32, mean, winstr / mean, 32.0 * winstr / mean, 32.0 * winstr * nsm / (ms * 1e-3) / 1e12, ms);
  }
  cudaFree(out); cudaFree(cyc);
}

// ---- wrap-semantics check ----
__global__ void sem_kernel(const unsigned* a, const unsigned* b, const unsigned* c, unsigned* r, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  r[0 * n + i] = __vadd2(a[i], b[i]);
  r[1 * n + i] = __vsub2(a[i], b[i]);
  r[2 * n + i] = __vmaxs2(a[i], b[i]);
  r[3 * n + i] = __viaddmax_s16x2(a[i], b[i], c[i]);
  r[4 * n + i] = __vimax3_s16x2(a[i], b[i], c[i]);
  r[5 * n + i] = __vneg2(a[i]);
}
static inline int16_t lo(unsigned x) { return (int16_t)(x & 0xffff); }
static inline int16_t hi(unsigned x) { return (int16_t)(x >> 16); }
static inline unsigned pk(int16_t l, int16_t h) { return (unsigned)(uint16_t)l | ((unsigned)(uint16_t)h << 16); }
static inline int16_t w16(int v) { return (int16_t)(uint16_t)(v & 0xffff); }
static inline int16_t mx(int16_t a, int16_t b) { return a > b ? a : b; }

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, clk_khz);
  // semantics
  {
    const int n = 1 << 20;
    unsigned *ha = (unsigned*)malloc(4 * n), *hb = (unsigned*)malloc(4 * n), *hc = (unsigned*)malloc(4 * n), *hr = (unsigned*)malloc(4 * n * 6);
    srand(1);
    for (int i = 0; i < n; i++) {
      ha[i] = ((unsigned)rand() << 16) ^ rand() ^ ((unsigned)rand() << 31);
      hb[i] = ((unsigned)rand() << 16) ^ rand() ^ ((unsigned)rand() << 31);
      hc[i] = ((unsigned)rand() << 16) ^ rand() ^ ((unsigned)rand() << 31);
      if (i % 7 == 0) { ha[i] = 0x7fff8000u; }
      if (i % 11 == 0) { hb[i] = 0x7fff8000u; }
    }
    unsigned *da, *db, *dc, *dr;
    CK(cudaMalloc(&da, 4 * n)); CK(cudaMalloc(&db, 4 * n)); CK(cudaMalloc(&dc, 4 * n)); CK(cudaMalloc(&dr, 4 * n * 6));
    CK(cudaMemcpy(da, ha, 4 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, 4 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dc, hc, 4 * n, cudaMemcpyHostToDevice));
    sem_kernel<<<n / 256, 256>>>(da, db, dc, dr, n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hr, dr, 4 * n * 6, cudaMemcpyDeviceToHost));
    long bad[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
      unsigned a = ha[i], b = hb[i], c = hc[i];
      unsigned e0 = pk(w16(lo(a) + lo(b)), w16(hi(a) + hi(b)));
      unsigned e1 = pk(w16(lo(a) - lo(b)), w16(hi(a) - hi(b)));
      unsigned e2 = pk(mx(lo(a), lo(b)), mx(hi(a), hi(b)));
      unsigned e3 = pk(mx(w16(lo(a) + lo(b)), lo(c)), mx(w16(hi(a) + hi(b)), hi(c)));
      unsigned e4 = pk(mx(mx(lo(a), lo(b)), lo(c)), mx(mx(hi(a), hi(b)), hi(c)));
      unsigned e5 = pk(w16(-lo(a)), w16(-hi(a)));
      bad[0] += hr[0 * n + i] != e0; bad[1] += hr[1 * n + i] != e1; bad[2] += hr[2 * n + i] != e2;
      bad[3] += hr[3 * n + i] != e3; bad[4] += hr[4 * n + i] != e4; bad[5] += hr[5 * n + i] != e5;
    }
    printf("WRAP-SEMANTICS mismatches vs wrapping-int16 model over %d random words: vadd2=%ld vsub2=%ld vmaxs2=%ld viaddmax_s16x2=%ld vimax3_s16x2=%ld vneg2=%ld\n",
           n, bad[0], bad[1], bad[2], bad[3], bad[4], bad[5]);
  }
  int nsm = p.multiProcessorCount;
  double g = clk_khz * 1e-6;
  run<OP_VADD2>(nsm, g); run<OP_VMAX2>(nsm, g); run<OP_VIADDMAX2>(nsm, g); run<OP_VIMAX3_2>(nsm, g);
  run<OP_IADD3>(nsm, g); run<OP_LOP3>(nsm, g); run<OP_IMAD>(nsm, g); run<OP_MIX_VADD_IMAD>(nsm, g);
  run<OP_MIX_VADD_VMAX>(nsm, g); run<OP_VIADDMAX32>(nsm, g); run<OP_MIX_VIADDMAX2_IMAD>(nsm, g); run<OP_MIX_VADD2_FFMA>(nsm, g); run<OP_PRMT>(nsm, g);
  run<OP_MIX_VADD2_VIADDMAX2>(nsm, g); run<OP_MIX_VMAX2_IMAD>(nsm, g); run<OP_MIX_VADD2_VIMAX3>(nsm, g); run<OP_MIX_1VADD_2VIADDMAX>(nsm, g); run<OP_MIX_VADD2_VMAX2_INDEP>(nsm, g); run<OP_MIX_VADD2_LOP3>(nsm, g); run<OP_MIX_VIADDMAX2_LOP3>(nsm, g);
  return 0;
}
