// Times scan_kernel<2> (the production kernel, included from the source tree) on synthetic streams: cycles per trellis step
// for 1 group, 2 groups (one block), and a machine-filling number of groups. Built with -DPROBE_VARIANT=n to test ideas.
#define SCAN_PROBE 1
#include <cstdio>
#include <vector>
#include "../../srsran_4g_b200/csrc/turbo_kernels.cuh"
using namespace srsb200;
int main()
{
  const uint32_t K = 6144, R = ((K + 3 + W - 1) / W) * W;
  const int maxg = 256;
  size_t wsg = ((group_ws_words(R) * 4 + 255) / 256) * 256;
  uint8_t* ws; cudaMalloc(&ws, wsg * maxg); cudaMemset(ws, 1, wsg * maxg);
  std::vector<Group> hg(maxg);
  for (int g = 0; g < maxg; g++) { hg[g].K = K; hg[g].R = R; hg[g].kidx = 187; hg[g].crc_kind = 2; hg[g].wpj = WPJ; hg[g].ws_off = wsg * g; for (int i = 0; i < 64; i++) hg[g].cb[i] = g * 64 + i; }
  Group* dg; cudaMalloc(&dg, sizeof(Group) * maxg); cudaMemcpy(dg, hg.data(), sizeof(Group) * maxg, cudaMemcpyHostToDevice);
  uint8_t* act; cudaMalloc(&act, maxg); cudaMemset(act, 1, maxg);
  cudaFuncSetAttribute(scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ScanSmemT<2>)));
  cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ScanSmemT<2>)));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int prop_clock; cudaDeviceGetAttribute(&prop_clock, cudaDevAttrClockRate, 0);
  for (int ng : {1, 2, 64, 128}) {
    for (int rep = 0; rep < 3; rep++) scan_kernel<2><<<(ng + 1) / 2, 160, sizeof(ScanSmemT<2>)>>>(dg, ws, act, ng);
    cudaEventRecord(a);
    for (int rep = 0; rep < 5; rep++) scan_kernel<2><<<(ng + 1) / 2, 160, sizeof(ScanSmemT<2>)>>>(dg, ws, act, ng);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    long long hc[1024]; cudaMemcpyFromSymbol(hc, g_probe_cycles, sizeof(long long) * 1024);
    if (ng <= 64) printf("   group 0: beta waited %lld cycles in %lld blocking waits, alpha %lld cycles in %lld waits (of %d chunks)\n", hc[512], hc[768], hc[513], hc[769], (int)(K / W + 1));
    double cb = 0, ca = 0; for (int g = 0; g < ng; g++) { cb += hc[2 * g]; ca += hc[2 * g + 1]; }
    printf("   in-kernel clock64: beta %.1f cycles/step, alpha %.1f cycles/step\n", cb / ng / (K + 3), ca / ng / K);
    printf("scan_kernel<2> groups=%3d : %.1f us per launch = %.1f cycles per step (at %.0f MHz nominal)  err=%s\n", ng, ms * 1e3, ms * 1e-3 * prop_clock * 1e3 / (K + 3), prop_clock / 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
