// cycles per step of the production beta_block / alpha_block code on a static shared-memory chunk (no copies, no barriers)
#include <cstdio>
#include "../../srsran_4g_b200/csrc/turbo_kernels.cuh"
using namespace srsb200;
template <int VAR>
__global__ void k(uint32_t* out, long long* cyc, uint32_t* ck, int iters)
{
  __shared__ ScanStageT<2> st;
  __shared__ uint32_t ckb[4][W / CKB][LANES][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * W * LANES; i += blockDim.x) (&st.s[0][0][0])[i] = i * 2654435761u;
  __syncthreads();
  uint32_t b[8];
  for (int i = 0; i < 8; i++) b[i] = threadIdx.x * 7 + i * 3;
  uint32_t* myck = ck + (size_t)(blockIdx.x * 4 + wid) * 1024 * 256;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (VAR == 3) {  // production beta block, row offset varies with the iteration so the loads cannot be hoisted
      beta_block<2, 16>(st, (it * 16) & 48, b, ckb[wid], lane);
    } else if (VAR == 4) {
      beta_block<2, 8>(st, (it * 8) & 56, b, ckb[wid], lane);
    } else if (VAR == 0) {  // production beta blocks over a 64-row chunk
      for (int top = 64; top > 0; top -= 16) beta_block<2, 16>(st, top - 16, b, ckb[wid], lane);
    } else if (VAR == 1) {  // alpha blocks
      for (int k0 = 0; k0 < 64; k0 += 16) alpha_block<2, 16>(st, k0, (it & 7) * 64 + k0 + 8, b, myck, lane);
    } else if (VAR == 2) {  // 8-step blocks
      for (int top = 64; top > 0; top -= 8) beta_block<2, 8>(st, top - 8, b, ckb[wid], lane);
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VAR> void run(const char* what, int warps)
{
  uint32_t *out, *ck; long long* cyc;
  cudaMalloc(&out, 4 * 148 * 128); cudaMalloc(&cyc, 8 * 148); cudaMalloc(&ck, (size_t)148 * 4 * 1024 * 256 * 4);
  k<VAR><<<148, 32 * warps>>>(out, cyc, ck, 10);
  k<VAR><<<148, 32 * warps>>>(out, cyc, ck, 2000);
  long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  double m = 0; for (int i = 0; i < 148; i++) m += h[i]; m /= 148;
  const double steps = (VAR == 3) ? 16 : (VAR == 4 ? 8 : 64);
  printf("%-46s warps/block=%d : %.1f cycles per step  (%s)\n", what, warps, m / (2000.0 * steps), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc); cudaFree(ck);
}
int main()
{
  run<0>("beta_block<16> x4 per 64-row chunk", 4); run<0>("beta_block<16> x4 per 64-row chunk", 1);
  run<1>("alpha_block<16> x4", 4); run<2>("beta_block<8> x8", 4);
  run<3>("beta_block<16>, varying rows (no hoisting)", 4); run<4>("beta_block<8>, varying rows (no hoisting)", 4);
  return 0;
}
