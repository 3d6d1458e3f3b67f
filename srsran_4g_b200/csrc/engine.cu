/*
 * engine.cu - host side of libsrsran_b200.so: device tables, batch planning, kernel launches and the C ABI declared in
 * include/srsran_b200.h. C++ because the reference's host side is compiled code (C); no torch types anywhere.
 * There is no CPU fallback: without a usable CUDA device every compute entry point returns SRSB200_ERROR_NO_DEVICE.
 */
#include "../../include/srsran_b200.h"
#include "../../include/lte_qpp_params.h"
#include "turbo_kernels.cuh"
#include "rm_kernels.cuh"
#include "tx_kernels.cuh"
#include "win8_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace srsb200;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* srsb200_last_error(void) { return g_err; }

#define CUDA_TRY(expr)                                                                                                 \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess) {                                                                                           \
      cudaGetLastError();                                                                                              \
      return fail(SRSB200_ERROR, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);          \
    }                                                                                                                  \
  } while (0)

// ------------------------------------------------------------------ host-side metadata
extern "C" int srsb200_cbsize(uint32_t index)
{
  return index < LTE_NOF_CB_SIZES ? (int)lte_qpp_params[index].K : SRSB200_ERROR;
}
extern "C" int srsb200_cbindex(uint32_t long_cb)
{
  // first table entry >= long_cb (srsran_cbsegm_cbindex, cbsegm.c:119-130)
  for (int i = 0; i < LTE_NOF_CB_SIZES; i++)
    if (lte_qpp_params[i].K >= long_cb) return i;
  return SRSB200_ERROR;
}
static int cbindex_exact(uint32_t K)
{
  // direct K -> index table (a 16384-block submission validates every block: a linear search per block costs ~1 ms)
  static const std::vector<int16_t> lut = [] {
    std::vector<int16_t> t(SRSB200_MAX_K + 1, (int16_t)-1);
    for (int i = 0; i < LTE_NOF_CB_SIZES; i++) t[lte_qpp_params[i].K] = (int16_t)i;
    return t;
  }();
  return K <= SRSB200_MAX_K ? lut[K] : -1;
}
extern "C" uint32_t srsb200_tdec_autoimp_get_subblocks(uint32_t) { return 0; }

extern "C" int srsb200_cbsegm(uint32_t tbs, uint32_t o[8])
{
  // 36.212 5.1.2 as realised by srsran_cbsegm (cbsegm.c:62-117)
  memset(o, 0, 8 * sizeof(uint32_t));
  if (tbs == 0) return SRSB200_SUCCESS;
  const uint32_t Z = SRSB200_MAX_K;
  uint32_t       B = tbs + 24, C = 1, Bp = B;
  if (B > Z) {
    C  = (B + (Z - 24) - 1) / (Z - 24);
    Bp = B + 24 * C;
  }
  int i1 = srsb200_cbindex((Bp - 1) / C + 1);
  if (i1 < 0) return SRSB200_ERROR;
  uint32_t K1 = lte_qpp_params[i1].K;
  o[1] = C; o[2] = K1; o[4] = (uint32_t)i1;
  if (C == 1) {
    o[6] = 1;
  } else {
    uint32_t K2 = lte_qpp_params[i1 - 1].K;
    o[3] = K2; o[5] = (uint32_t)(i1 - 1);
    o[7] = (C * K1 - Bp) / (K1 - K2);
    o[6] = C - o[7];
  }
  o[0] = o[6] * o[2] + o[7] * o[3] - Bp;
  return SRSB200_SUCCESS;
}

// ------------------------------------------------------------------ tables
static void qpp_tables(int kidx, std::vector<uint16_t>& fwd, std::vector<uint16_t>& rev)
{
  const uint32_t K = lte_qpp_params[kidx].K;
  const uint64_t f1 = lte_qpp_params[kidx].f1, f2 = lte_qpp_params[kidx].f2;
  fwd.resize(K);
  rev.resize(K);
  for (uint64_t i = 0; i < K; i++) {
    uint32_t p = (uint32_t)((f1 * i + f2 * i * i) % K);
    fwd[i]     = (uint16_t)p;
    rev[p]     = (uint16_t)i;
  }
}

// x^(m+24) mod g for m = 0..MAX_K-1: contribution of a set bit m positions before the end of the message to the
// MSB-first, zero-init CRC of srsran_crc_checksum_byte (crc.c:147-161)
static void crc_position_words(uint32_t poly, std::vector<uint32_t>& r)
{
  r.resize(SRSB200_MAX_K);
  uint32_t v = 1;  // x^0
  // advance to x^24
  auto mulx = [&](uint32_t a) {
    a <<= 1;
    if (a & 0x1000000u) a ^= poly;
    return a & 0xffffffu;
  };
  for (int i = 0; i < 24; i++) v = mulx(v);
  for (uint32_t m = 0; m < SRSB200_MAX_K; m++) {
    r[m] = v;
    v    = mulx(v);
  }
}

static const uint8_t RM_COLPERM[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                       1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

// 36.212 5.1.4.1 sub-block interleaver + circular buffer walk: transmitted bit n of redundancy version rv is element
// T[n] of the natural coded stream (what srsran_rm_turbo_gentable_receive builds, rm_turbo.c:175-248)
static void rm_table_host(uint32_t cb_idx, uint32_t rv, std::vector<uint16_t>& T)
{
  const uint32_t K = lte_qpp_params[cb_idx].K, D = K + 4, Rr = (D - 1) / 32 + 1, Kp = 32 * Rr, Nd = Kp - D, Ncb = 3 * Kp, L = 3 * D;
  std::vector<int32_t> w(Ncb, -1);
  for (uint32_t i = 0; i < D; i++) {
    uint32_t y = Nd + i, v = RM_COLPERM[y % 32] * Rr + y / 32;
    w[v]          = (int32_t)(3 * i);
    w[Kp + 2 * v] = (int32_t)(3 * i + 1);
    uint32_t t = (y + Kp - 1) % Kp, v2 = RM_COLPERM[t % 32] * Rr + t / 32;
    w[Kp + 2 * v2 + 1] = (int32_t)(3 * i + 2);
  }
  T.resize(L);
  uint32_t k0 = Rr * (24 * rv + 2), n = 0;
  for (uint32_t j = 0; n < L; j++) {
    int32_t src = w[(k0 + j) % Ncb];
    if (src >= 0) T[n++] = (uint16_t)src;
  }
}

extern "C" int srsb200_rm_table(uint32_t cb_idx, uint32_t rv_idx, uint16_t* table)
{
  if (cb_idx >= LTE_NOF_CB_SIZES || rv_idx >= 4 || !table) return SRSB200_ERROR_INVALID_INPUTS;
  std::vector<uint16_t> T;
  rm_table_host(cb_idx, rv_idx, T);
  memcpy(table, T.data(), T.size() * sizeof(uint16_t));
  return SRSB200_SUCCESS;
}

// ------------------------------------------------------------------ engine
struct srsb200_engine {
  int          device = 0;
  cudaStream_t stream = nullptr;
  uint64_t     launches = 0;
  std::mutex   mtx;

  // per-K device tables
  KTable              h_ktab[LTE_NOF_CB_SIZES];
  bool                have_k[LTE_NOF_CB_SIZES];
  KTable*             d_ktab = nullptr;
  std::vector<void*>  owned;  // device allocations freed at destroy
  std::vector<uint32_t> crc_words[3];

  // x^(m+24) mod g24A for the transport-block CRC kernel, grown on demand
  uint32_t* d_tb_crc_words = nullptr;
  uint32_t  tb_crc_words_n = 0;

  // rate-matching tables [cb_idx][rv] (device, uint16[3K+12]) built lazily
  uint16_t* d_rm[LTE_NOF_CB_SIZES][4];      // T    (transmit side: e[n] = coded[T[n mod L]])
  uint16_t* d_rm_inv[LTE_NOF_CB_SIZES][4];  // Tinv (receive side: soft-buffer position p collects e[Tinv[p] + kL])

  // scratch for the host-pointer APIs (grown on demand)
  void*  d_scratch[20]   = {nullptr};  // 0-7 receive side, 8-11 transmit side, 12-15 UL-SCH de-interleaver / demodulation, 16 reset list
  size_t scratch_cap[20] = {0};
  uint32_t* d_crc24b_words = nullptr;  // x^(m+24) mod g24B, m < 6144 (tx_cb_kernel)
  uint32_t* d_gold = nullptr;          // jump matrices of the scrambling LFSRs [2][GOLD_POWERS][32] (rm_rx_kernel)
  uint32_t  h_gold[2][GOLD_POWERS][32];

  // pinned staging arenas for pageable caller buffers (0: host->device, 1: device->host), see Stager
  void*  h_stage[2]     = {nullptr, nullptr};
  size_t h_stage_cap[2] = {0, 0};

  // sub-batch streams (see launch_plan)
  static const int MAX_SUB = 16;
  int          n_sub = 8;      // ranges of a host-pointer submission (copy/compute overlap)
  int          n_sub_dev = 4;  // ranges of a device-resident submission (their launch chains overlap a little: +3-4 %)
  // Device-resident submissions run on one of two LANES (one half of sub[] each; a plan is bound to a lane when it is built)
  // and are joined back into `stream` lazily (join_pending): the latency-bound last half-iterations of one submission
  // then overlap the bandwidth-bound first ones of the next submission of ANOTHER plan (+11 % on the bench workload).
  int          n_lanes = 2;    // env SRSB200_LANES (1..4): MAX_SUB / n_lanes streams each
  int          next_lane = 0;
  bool         pending[MAX_SUB] = {false};
  cudaStream_t sub[MAX_SUB] = {nullptr};
  cudaEvent_t  ev_fork = nullptr, ev_join[MAX_SUB] = {nullptr};

  std::vector<int16_t*> pending_zero;  // soft-buffer mirrors reset since the last submission (zeroed in front of the next use)
  int fail_alloc_countdown = 0;  // > 0: the n-th ensure_scratch call from now fails (tests of the error paths)
  // optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg)
  bool profiling = false;
  struct ProfEv { cudaEvent_t a, b; int kind; };
  std::vector<ProfEv> prof;

  // device-resident HARQ soft buffers (srsb200_softbuffer_*): host buffer_f pointer -> device copy
  bool softbuffer_resident = false;
  std::unordered_map<const void*, int16_t*> softslots;
  std::vector<int16_t*> softslot_free;
  std::vector<void*>    softslot_chunks;

  // last plan built by srsb200_decode_tb_batch and the shape it was built for
  struct srsb200_plan* tb_plan = nullptr;
  std::vector<uint32_t> tb_K, tb_olen;
  std::vector<uint8_t>  tb_kind;
  std::vector<uint64_t> tb_boff, tb_ooff;

  // 8-bit LLR mode (win8.inc): tables + the last plans of srsb200_tdec_batch8 / the 8-bit transport-block path
  void* w8 = nullptr;
  struct srsb200_plan8* cached_plan8 = nullptr;
  struct srsb200_plan8* tb_plan8 = nullptr;
  std::vector<uint32_t> tb8_K, tb8_olen;
  std::vector<uint8_t>  tb8_kind;
  std::vector<uint64_t> tb8_boff, tb8_ooff;

  // last plan built by srsb200_tdec_batch, reused when the next submission has the same shape
  struct srsb200_plan* cached_plan = nullptr;
  std::vector<uint32_t> cached_K;
  std::vector<uint8_t>  cached_kind;
  std::vector<uint64_t> cached_loff, cached_ooff;
};

static int join_pending(srsb200_engine* e);
static void plan8_destroy(struct srsb200_plan8* p);
static void w8_engine_free(srsb200_engine* e);

struct ProfScope {
  srsb200_engine* e; int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(srsb200_engine* e_, int k, cudaStream_t s = nullptr) : e(e_), kind(k), st(s ? s : e_->stream)
  {
    if (e->profiling) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope()
  {
    if (a) {
      cudaEventRecord(b, st);
      e->prof.push_back({a, b, kind});
    }
  }
};

static int ensure_scratch(srsb200_engine* e, int slot, size_t bytes, void** out)
{
  // fault injection (srsb200_engine_inject_alloc_failure / SRSB200_FAIL_ALLOC): the n-th scratch request from now on fails
  if (e->fail_alloc_countdown > 0 && --e->fail_alloc_countdown == 0) return fail(SRSB200_ERROR, "injected allocation failure (scratch slot %d)", slot);
  if (e->scratch_cap[slot] < bytes) {
    if (e->d_scratch[slot]) cudaFree(e->d_scratch[slot]);
    e->d_scratch[slot]   = nullptr;
    e->scratch_cap[slot] = 0;
    size_t cap = bytes + bytes / 4 + 4096;
    CUDA_TRY(cudaMalloc(&e->d_scratch[slot], cap));
    e->scratch_cap[slot] = cap;
  }
  *out = e->d_scratch[slot];
  return 0;
}

/*
 * Copies between caller memory and the device. cudaMemcpyAsync on PAGEABLE memory is synchronous and costs ~10-30 us per
 * call (the driver stages it) - a transport-block submission makes hundreds of them (e-bits per TB, soft buffer per code
 * block). The Stager sends pinned/registered caller buffers straight to the DMA engine and routes pageable ones through
 * an engine-owned pinned arena: host memcpy of buffer i+1 overlaps the DMA of buffer i, and device->host results are
 * copied out of the arena after the one final synchronise.
 */
struct Stager {
  srsb200_engine* e;
  size_t used[2] = {0, 0};
  struct Out { void* dst; const void* src; size_t bytes; };
  std::vector<Out>     outs;
  std::vector<CopyJob> q;  // transfers queued since the last flush (device-visible addresses on both sides)
  explicit Stager(srsb200_engine* e_) : e(e_) {}
  // both arenas must be sized before the first copy (they cannot move while copies are in flight); max_jobs = upper bound
  // of the transfers this call will queue (their job lists live in the host->device arena)
  int reserve(size_t in_bytes, size_t out_bytes, size_t max_jobs = 64)
  {
    const size_t want[2] = {in_bytes + (max_jobs + 64) * sizeof(CopyJob) + 4096, out_bytes};
    for (int i = 0; i < 2; i++) {
      if (e->h_stage_cap[i] >= want[i]) continue;
      if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
      e->h_stage[i]     = nullptr;
      e->h_stage_cap[i] = 0;
      const size_t cap = want[i] + want[i] / 4 + 65536;
      CUDA_TRY(cudaHostAlloc(&e->h_stage[i], cap, cudaHostAllocDefault));
      e->h_stage_cap[i] = cap;
    }
    return 0;
  }
  // page-locked (cudaHostAlloc / cudaHostRegister) memory -> the address the GPU uses for it, else nullptr
  static void* device_view(const void* p)
  {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
  }
  void* arena_take(int which, size_t bytes)
  {
    const size_t off = (used[which] + 63) & ~(size_t)63;
    if (off + bytes > e->h_stage_cap[which]) return nullptr;
    used[which] = off + bytes;
    return (uint8_t*)e->h_stage[which] + off;
  }
  cudaError_t h2d(void* dst, const void* src, size_t bytes, cudaStream_t st)
  {
    if (bytes == 0) return cudaSuccess;
    const void* dv = device_view(src);
    if (!dv) {
      void* s = arena_take(0, bytes);
      if (!s) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);  // arena exhausted: the slow way
      memcpy(s, src, bytes);
      dv = s;  // cudaHostAlloc memory: same address on the device (unified virtual addressing)
    }
    q.push_back({dv, dst, bytes});
    return cudaSuccess;
  }
  cudaError_t d2h(void* dst, const void* src, size_t bytes, cudaStream_t st)
  {
    if (bytes == 0) return cudaSuccess;
    void* dv = device_view(dst);
    if (!dv) {
      void* s = arena_take(1, bytes);
      if (!s) {
        cudaError_t ce = flush(st);
        return ce != cudaSuccess ? ce : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
      }
      outs.push_back({dst, s, bytes});
      dv = s;
    }
    q.push_back({src, dv, bytes});
    return cudaSuccess;
  }
  // issue everything queued so far, in stream order: call before a kernel that consumes uploaded data / before the final
  // synchronise. Neighbouring transfers that are contiguous on both sides are merged (buffers staged through the arena
  // always are); what is then still large goes to the copy engines (a kernel pulling megabytes across PCIe would hold
  // SM slots for the whole transfer), the many small ones - job lists, decoded bytes, flags - become ONE
  // gather_copy_kernel launch instead of one driver call each.
  cudaError_t flush(cudaStream_t st)
  {
    if (q.empty()) return cudaSuccess;
    std::vector<CopyJob> m;
    for (auto& j : q) {
      if (!m.empty() && (const uint8_t*)m.back().src + m.back().bytes == (const uint8_t*)j.src && (uint8_t*)m.back().dst + m.back().bytes == (uint8_t*)j.dst)
        m.back().bytes += j.bytes;
      else
        m.push_back(j);
    }
    q.clear();
    cudaError_t ce = cudaSuccess;
    size_t      nsmall = 0;
    for (auto& j : m) nsmall += j.bytes < BIG;
    CopyJob* list = nsmall > 2 ? (CopyJob*)arena_take(0, nsmall * sizeof(CopyJob)) : nullptr;
    size_t   k    = 0;
    uint64_t maxb = 0;
    for (auto& j : m) {
      if (j.bytes >= BIG || !list) {
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyDefault, st);
      } else {
        list[k++] = j;
        maxb      = std::max<uint64_t>(maxb, j.bytes);
      }
    }
    for (size_t k0 = 0; k0 < k && ce == cudaSuccess; k0 += MAX_GRID_Y) {  // gridDim.y <= 65535
      const size_t kn = std::min<size_t>(MAX_GRID_Y, k - k0);
      gather_copy_kernel<<<dim3((unsigned)std::max<uint64_t>(1, (maxb + 16383) / 16384), (unsigned)kn), 256, 0, st>>>(list + k0);
      e->launches++;
      ce = cudaGetLastError();
    }
    return ce;
  }
  static constexpr uint64_t BIG = 32768;
  static constexpr size_t   MAX_GRID_Y = 65535;
  // error path: forget what has not been issued (the caller drains the stream)
  void drop()
  {
    q.clear();
    outs.clear();
  }
  // after the stream has been synchronised
  void finish()
  {
    for (auto& o : outs) memcpy(o.dst, o.src, o.bytes);
    outs.clear();
  }
};

static int ensure_ktable(srsb200_engine* e, int kidx)
{
  if (e->have_k[kidx]) return 0;
  const uint32_t K = lte_qpp_params[kidx].K;
  const uint32_t R = ((K + 3 + W - 1) / W) * W;
  std::vector<uint16_t> fwd, rev;
  qpp_tables(kidx, fwd, rev);
  KTable kt;
  memset(&kt, 0, sizeof(kt));
  {
    std::vector<uint32_t> r1(R, 0u), r2(R, 0u);
    for (uint32_t i = 0; i < K; i++) {
      r1[i] = rev[i];
      r2[i] = fwd[i];
    }
    uint32_t *d1, *d2;
    CUDA_TRY(cudaMalloc(&d1, R * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d2, R * sizeof(uint32_t)));
    e->owned.push_back(d1);
    e->owned.push_back(d2);
    CUDA_TRY(cudaMemcpyAsync(d1, r1.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d2, r2.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    kt.row1 = d1;
    kt.row2 = d2;
  }
  for (int kind = 1; kind < 3; kind++) {
    // nibble tables: contribution of every 4-bit group of decisions of a 16-step window to the CRC
    const uint32_t nw = R / 16;
    std::vector<uint32_t> n1((size_t)nw * 64, 0u), n2((size_t)nw * 64, 0u);
    for (uint32_t w = 0; w < nw; w++)
      for (uint32_t j = 0; j < 4; j++)
        for (uint32_t v = 1; v < 16; v++) {
          uint32_t x1 = 0, x2 = 0;
          for (uint32_t b = 0; b < 4; b++) {
            const uint32_t i = 16 * w + 4 * j + b;
            if (((v >> b) & 1u) && i < K) {
              x1 ^= e->crc_words[kind][K - 1 - i];
              x2 ^= e->crc_words[kind][K - 1 - fwd[i]];
            }
          }
          n1[((size_t)w * 4 + j) * 16 + v] = x1;
          n2[((size_t)w * 4 + j) * 16 + v] = x2;
        }
    uint32_t *d1, *d2;
    CUDA_TRY(cudaMalloc(&d1, n1.size() * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d2, n2.size() * sizeof(uint32_t)));
    e->owned.push_back(d1);
    e->owned.push_back(d2);
    CUDA_TRY(cudaMemcpyAsync(d1, n1.data(), n1.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d2, n2.data(), n2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));  // the host vectors die at the end of this scope
    kt.nib1[kind] = d1;
    kt.nib2[kind] = d2;
  }
  kt.nib1[0] = kt.nib1[2];  // CRC_NONE: the result is ignored, any table will do
  kt.nib2[0] = kt.nib2[2];
  uint16_t* drev;
  CUDA_TRY(cudaMalloc(&drev, K * sizeof(uint16_t)));
  e->owned.push_back(drev);
  CUDA_TRY(cudaMemcpyAsync(drev, rev.data(), K * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
  kt.rev = drev;
  e->h_ktab[kidx] = kt;
  CUDA_TRY(cudaMemcpyAsync(e->d_ktab + kidx, &e->h_ktab[kidx], sizeof(KTable), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->have_k[kidx] = true;
  return 0;
}

static int ensure_rm_table(srsb200_engine* e, uint32_t cb_idx, uint32_t rv)
{
  if (e->d_rm[cb_idx][rv]) return 0;
  std::vector<uint16_t> T;
  rm_table_host(cb_idx, rv, T);
  const size_t Lp = (T.size() + 7) & ~(size_t)7;  // the inverse table starts 16-byte aligned (rm_rx_kernel reads it 128 bits at a time)
  std::vector<uint16_t> both(Lp + T.size() + 8, 0);
  for (size_t n = 0; n < T.size(); n++) {
    both[n]         = T[n];
    both[Lp + T[n]] = (uint16_t)n;  // inverse permutation
  }
  uint16_t* d;
  CUDA_TRY(cudaMalloc(&d, both.size() * sizeof(uint16_t)));
  e->owned.push_back(d);
  CUDA_TRY(cudaMemcpyAsync(d, both.data(), both.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->d_rm[cb_idx][rv]     = d;
  e->d_rm_inv[cb_idx][rv] = d + Lp;
  return 0;
}

// ------------------------------------------------------------------ device-resident soft buffers
static const size_t SOFTSLOT_ELEMS = 18688;  // SOFTBUFFER_SIZE (18600) rounded up to a multiple of 64 int16
static int softslot_get(srsb200_engine* e, const void* host, int16_t** out, bool upload_if_new)
{
  auto it = e->softslots.find(host);
  if (it != e->softslots.end()) {
    *out = it->second;
    return 0;
  }
  if (e->softslot_free.empty()) {
    const size_t per_chunk = 1024;
    void* chunk = nullptr;
    CUDA_TRY(cudaMalloc(&chunk, per_chunk * SOFTSLOT_ELEMS * sizeof(int16_t)));
    e->softslot_chunks.push_back(chunk);
    for (size_t i = 0; i < per_chunk; i++) e->softslot_free.push_back((int16_t*)chunk + (per_chunk - 1 - i) * SOFTSLOT_ELEMS);
  }
  int16_t* d = e->softslot_free.back();
  e->softslot_free.pop_back();
  e->softslots[host] = d;
  // first touch of a soft buffer that was never reset through srsb200_softbuffer_reset: adopt the host content
  if (upload_if_new) CUDA_TRY(cudaMemcpyAsync(d, host, SRSB200_SOFTBUFFER_SIZE * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
  *out = d;
  return 0;
}

// srsb200_softbuffer_reset only notes which mirrors to zero; one kernel zeroes them in front of the next operation that reads or
// writes a mirror (a transport-block submission, a copy back to the host, a release). A reset per transport block used to cost a
// launch and a stream synchronisation each - 64 of them per subframe in the multi-cell case, more host time than the decode.
static int flush_pending_zero(srsb200_engine* e)
{
  if (e->pending_zero.empty()) return 0;
  const uint32_t n = (uint32_t)e->pending_zero.size();
  void* d_list;
  if (ensure_scratch(e, 16, sizeof(int16_t*) * n, &d_list)) return SRSB200_ERROR;
  // (pageable source: cudaMemcpyAsync returns once the list has been staged, so the vector may be cleared right away)
  CUDA_TRY(cudaMemcpyAsync(d_list, e->pending_zero.data(), sizeof(int16_t*) * n, cudaMemcpyHostToDevice, e->stream));
  for (uint32_t c0 = 0; c0 < n; c0 += 65535u) {  // gridDim.y <= 65535
    zero_slots_kernel<<<dim3((SOFTSLOT_ELEMS / 8 + 255) / 256, std::min(65535u, n - c0)), 256, 0, e->stream>>>((int16_t* const*)d_list + c0,
                                                                                                           (uint32_t)(SOFTSLOT_ELEMS / 8));
    e->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  e->pending_zero.clear();
  return 0;
}

extern "C" int srsb200_softbuffer_set_resident(srsb200_engine_t* e, int resident)
{
  if (!e) return SRSB200_ERROR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(e->mtx);
  e->softbuffer_resident = resident != 0;
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_softbuffer_reset(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb)
{
  if (!e) return SRSB200_ERROR_NO_DEVICE;
  if (!buffer_f) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  for (uint32_t i = 0; i < nof_cb; i++) {
    int16_t* d = nullptr;
    if (softslot_get(e, buffer_f[i], &d, false)) return SRSB200_ERROR;
    e->pending_zero.push_back(d);
  }
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_softbuffer_sync_to_host(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb)
{
  if (!e) return SRSB200_ERROR_NO_DEVICE;
  if (!buffer_f) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  if (flush_pending_zero(e)) return SRSB200_ERROR;
  for (uint32_t i = 0; i < nof_cb; i++) {
    auto it = e->softslots.find(buffer_f[i]);
    if (it != e->softslots.end())
      CUDA_TRY(cudaMemcpyAsync(buffer_f[i], it->second, SRSB200_SOFTBUFFER_SIZE * sizeof(int16_t), cudaMemcpyDeviceToHost, e->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_softbuffer_release(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb)
{
  if (!e) return SRSB200_ERROR_NO_DEVICE;
  if (!buffer_f) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  // a released mirror may be handed to another soft buffer before the next submission: it must not be zeroed under its new owner
  for (uint32_t i = 0; i < nof_cb; i++) {
    auto it = e->softslots.find(buffer_f[i]);
    if (it != e->softslots.end()) {
      e->pending_zero.erase(std::remove(e->pending_zero.begin(), e->pending_zero.end(), it->second), e->pending_zero.end());
      e->softslot_free.push_back(it->second);
      e->softslots.erase(it);
    }
  }
  return SRSB200_SUCCESS;
}

static int ensure_tb_crc_words(srsb200_engine* e, uint32_t nbits)
{
  if (e->tb_crc_words_n >= nbits) return 0;
  uint32_t n = std::max(nbits, 160000u);
  std::vector<uint32_t> w(n);
  uint32_t v = 1;
  auto mulx = [](uint32_t a) {
    a <<= 1;
    if (a & 0x1000000u) a ^= 0x1864CFBu;
    return a & 0xffffffu;
  };
  for (int i = 0; i < 24; i++) v = mulx(v);
  for (uint32_t m = 0; m < n; m++) {
    w[m] = v;
    v    = mulx(v);
  }
  if (e->d_tb_crc_words) cudaFree(e->d_tb_crc_words);
  e->d_tb_crc_words = nullptr;
  e->tb_crc_words_n = 0;
  CUDA_TRY(cudaMalloc(&e->d_tb_crc_words, sizeof(uint32_t) * n));
  CUDA_TRY(cudaMemcpyAsync(e->d_tb_crc_words, w.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->tb_crc_words_n = n;
  return 0;
}

// ------------------------------------------------------------------ scrambling sequence (36.211 7.2), GF(2) jump-ahead
// An LFSR window w (bit b = x(n + b), 31 bits) advances by one position as w' = (w >> 1) | (parity(w & taps) << 30). That map
// is linear, so advancing by any offset is a product of the matrices A^(2^k); a matrix is kept as 31 row masks.
static uint32_t gold_matvec_h(const uint32_t* rows, uint32_t x)
{
  uint32_t y = 0;
  for (int r = 0; r < 31; r++) y |= (uint32_t)(__builtin_popcount(rows[r] & x) & 1) << r;
  return y;
}
static uint32_t gold_jump_h(const uint32_t (*jump)[32], uint32_t x, uint32_t off)
{
  for (int k = 0; off; k++, off >>= 1)
    if (off & 1u) x = gold_matvec_h(jump[k], x);
  return x;
}
static void gold_build(uint32_t h[2][GOLD_POWERS][32])
{
  const uint32_t taps[2] = {0x9u, 0xFu};  // x1: x(n) + x(n+3); x2: x(n) + x(n+1) + x(n+2) + x(n+3)
  for (int q = 0; q < 2; q++) {
    for (int r = 0; r < 30; r++) h[q][0][r] = 1u << (r + 1);
    h[q][0][30] = taps[q];
    h[q][0][31] = 0;
    for (int k = 1; k < GOLD_POWERS; k++) {
      for (int r = 0; r < 31; r++) {  // row r of A^(2^k) = row r of A^(2^(k-1)) times A^(2^(k-1))
        uint32_t acc = 0, m = h[q][k - 1][r];
        for (int j = 0; j < 31; j++)
          if ((m >> j) & 1u) acc ^= h[q][k - 1][j];
        h[q][k][r] = acc;
      }
      h[q][k][31] = 0;
    }
  }
}

extern "C" int srsb200_engine_create(srsb200_engine_t** out, int device)
{
  if (!out) return fail(SRSB200_ERROR_INVALID_INPUTS, "null engine pointer");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SRSB200_ERROR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return fail(SRSB200_ERROR_NO_DEVICE, "cudaGetDevice failed");
  }
  if (device >= ndev) return fail(SRSB200_ERROR_INVALID_INPUTS, "device %d out of range (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  srsb200_engine* e = new srsb200_engine();
  e->device         = device;
  memset(e->have_k, 0, sizeof(e->have_k));
  memset(e->d_rm, 0, sizeof(e->d_rm));
  memset(e->d_rm_inv, 0, sizeof(e->d_rm_inv));
  memset(e->h_ktab, 0, sizeof(e->h_ktab));
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  if (const char* env = getenv("SRSB200_FAIL_ALLOC")) e->fail_alloc_countdown = std::max(0, atoi(env));
  if (const char* env = getenv("SRSB200_SUBBATCHES")) e->n_sub = std::max(1, std::min((int)srsb200_engine::MAX_SUB, atoi(env)));
  if (const char* env = getenv("SRSB200_LANES")) e->n_lanes = std::max(1, std::min(4, atoi(env)));
  if (const char* env = getenv("SRSB200_SUBBATCHES_DEV")) e->n_sub_dev = std::max(1, std::min((int)srsb200_engine::MAX_SUB, atoi(env)));
  for (int i = 0; i < srsb200_engine::MAX_SUB; i++) {
    CUDA_TRY(cudaStreamCreateWithFlags(&e->sub[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaMalloc(&e->d_ktab, sizeof(KTable) * LTE_NOF_CB_SIZES));
  CUDA_TRY(cudaMemset(e->d_ktab, 0, sizeof(KTable) * LTE_NOF_CB_SIZES));
  crc_position_words(0x1864CFBu, e->crc_words[SRSB200_CRC_24A]);
  crc_position_words(0x1800063u, e->crc_words[SRSB200_CRC_24B]);
  CUDA_TRY(cudaFuncSetAttribute(scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmemT<0>)));
  CUDA_TRY(cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmemT<1>)));
  CUDA_TRY(cudaFuncSetAttribute(scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmemT<2>)));
  CUDA_TRY(cudaFuncSetAttribute(rm_rx_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((RM_SMEM_ELEMS + 8) * sizeof(int16_t))));
  CUDA_TRY(cudaFuncSetAttribute(rm_rx_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((RM_SMEM_ELEMS + 8) * sizeof(int16_t))));
  gold_build(e->h_gold);
  CUDA_TRY(cudaMalloc(&e->d_gold, sizeof(e->h_gold)));
  CUDA_TRY(cudaMemcpy(e->d_gold, e->h_gold, sizeof(e->h_gold), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaFuncSetAttribute(emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)emit_smem_bytes(((SRSB200_MAX_K + 3 + W - 1) / W) * W, SRSB200_MAX_K + 64)));
  CUDA_TRY(cudaFuncSetAttribute(job_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(JobWarpSmem))));
  CUDA_TRY(cudaFuncSetAttribute(job_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(JobWarpSmem))));
  CUDA_TRY(cudaFuncSetAttribute(job_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(JobWarpSmem))));
  *out = e;
  return SRSB200_SUCCESS;
}

extern "C" void srsb200_engine_destroy(srsb200_engine_t* e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  join_pending(e);
  cudaStreamSynchronize(e->stream);
  if (e->cached_plan) srsb200_plan_destroy(e->cached_plan);
  if (e->tb_plan) srsb200_plan_destroy(e->tb_plan);
  plan8_destroy(e->cached_plan8);
  plan8_destroy(e->tb_plan8);
  w8_engine_free(e);
  for (void* p : e->softslot_chunks) cudaFree(p);
  for (void* p : e->owned) cudaFree(p);
  for (int i = 0; i < 16; i++)
    if (e->d_scratch[i]) cudaFree(e->d_scratch[i]);
  cudaFree(e->d_ktab);
  cudaFree(e->d_gold);
  for (int i = 0; i < 2; i++)
    if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
  if (e->d_tb_crc_words) cudaFree(e->d_tb_crc_words);
  for (int i = 0; i < srsb200_engine::MAX_SUB; i++) {
    if (e->sub[i]) cudaStreamDestroy(e->sub[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int srsb200_engine_profile(srsb200_engine_t* e, int enable)
{
  if (!e) return SRSB200_ERROR_INVALID_INPUTS;
  e->profiling = enable != 0;
  return SRSB200_SUCCESS;
}
// ms[kind] += elapsed, cnt[kind] += launches; kinds: 0 extract, 2 emit, 3 rate-dematch, 4 tb-crc, 5 scan, 6 job, 7 tb-encode.
// Synchronises. While profiling is on, decodes run as a single chain (no sub-batch overlap).
extern "C" int srsb200_engine_profile_read_kinds(srsb200_engine_t* e, double* ms, uint64_t* cnt, uint32_t n_kinds)
{
  if (!e || !ms || !cnt) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (uint32_t i = 0; i < n_kinds; i++) { ms[i] = 0; cnt[i] = 0; }
  for (auto& p : e->prof) {
    float t = 0;
    cudaEventElapsedTime(&t, p.a, p.b);
    // kinds the caller's arrays have no room for are folded into the kind they belonged to before they were split off
    uint32_t k = (uint32_t)p.kind;
    if (k >= n_kinds) k = (k == 10) ? 0u : 3u;
    if (k < n_kinds) { ms[k] += t; cnt[k]++; }
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  e->prof.clear();
  return SRSB200_SUCCESS;
}
extern "C" int srsb200_engine_profile_read(srsb200_engine_t* e, double ms[8], uint64_t cnt[8]) { return srsb200_engine_profile_read_kinds(e, ms, cnt, 8); }

extern "C" int srsb200_engine_set_subbatches(srsb200_engine_t* e, int n)
{
  if (!e || n < 1 || n > srsb200_engine::MAX_SUB) return SRSB200_ERROR_INVALID_INPUTS;
  e->n_sub = n;
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_engine_inject_alloc_failure(srsb200_engine_t* e, int nth)
{
  if (!e || nth < 0) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  e->fail_alloc_countdown = nth;
  return SRSB200_SUCCESS;
}

extern "C" uint64_t srsb200_engine_launch_count(const srsb200_engine_t* e) { return e ? e->launches : 0; }
extern "C" void*    srsb200_engine_stream(const srsb200_engine_t* e) { return e ? (void*)e->stream : nullptr; }
extern "C" void* srsb200_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void srsb200_host_free(void* p)
{
  if (p) cudaFreeHost(p);
}
extern "C" int srsb200_host_register(void* p, size_t bytes)
{
  if (!p || !bytes) return SRSB200_ERROR_INVALID_INPUTS;
  if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
    cudaGetLastError();
    return SRSB200_ERROR;
  }
  return SRSB200_SUCCESS;
}
extern "C" int srsb200_host_unregister(void* p)
{
  if (!p) return SRSB200_ERROR_INVALID_INPUTS;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return SRSB200_ERROR;
  }
  return SRSB200_SUCCESS;
}
extern "C" int      srsb200_engine_flush(srsb200_engine_t* e)
{
  if (!e) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  return join_pending(e) ? SRSB200_ERROR : SRSB200_SUCCESS;
}
extern "C" int      srsb200_engine_sync(srsb200_engine_t* e)
{
  if (!e) return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SRSB200_SUCCESS;
}

// ------------------------------------------------------------------ plans: code blocks -> groups
struct srsb200_plan {
  srsb200_engine* e = nullptr;
  uint32_t n_cb = 0, n_groups = 0, max_R = 0;
  std::vector<Group> h_groups;
  Group*    d_groups = nullptr;
  uint8_t*  d_ws = nullptr;
  uint64_t  ws_bytes = 0;
  uint64_t* d_llr_off = nullptr;
  uint64_t* d_out_off = nullptr;
  uint32_t* d_out_len = nullptr;  // optional [n_cb] bytes to emit per code block (nullptr: K/8)
  uint8_t*  d_cb_max_iter = nullptr;  // optional [n_cb] half-iteration limit per code block (use_cb_max_iter)
  bool      use_cb_max_iter = false;
  // decode state that lives across the launches of one decode (and across srsb200_tdec_iteration calls)
  uint32_t* d_crc_acc = nullptr;  // [n_cb] running CRC of the current half-iteration
  uint8_t*  d_done = nullptr;     // [n_cb]
  uint8_t*  d_active = nullptr;   // [n_groups]
  uint32_t* d_arrivals = nullptr; // [n_groups] job blocks of the group that have finished the current half-iteration
  // regrouping of the unfinished code blocks (turbo_kernels.cuh, regroup_plan_kernel): plans of one block size with enough groups
  // carry n_new spare group slots behind their own groups (table, workspace, activity flags, arrival counters)
  uint32_t  n_new = 0;
  int32_t*  d_home = nullptr;     // [n_cb] group (relative to the range's first) whose decision arrays hold the block's bits
  int32_t*  d_src = nullptr;      // [n_new][64] where each lane of a new group took its state from
  uint32_t* d_rg_state = nullptr; // [MAX_SUB] per running range: 0, or the regrouping point that regrouped it
  bool      uniform = false;
  int       lane = 0;            // which pair of sub-stream sets runs its device-resident submissions
  uint32_t  wpj = WPJ;           // windows per job warp: 16 for machine-filling batches (throughput), 8 otherwise (latency)
  // the launch chain of a small (single-range) decode as an instantiated CUDA graph, valid for exactly these arguments
  static const int N_GRAPHS = 4;  // a caller alternating a few buffer sets (double buffering, two codewords) re-uses, not re-captures
  struct GraphKey {
    const void *llr, *out, *noi, *ok, *mi;
    uint32_t max_iter, min_iter, start_iter;
    int early_stop, do_extract;
    bool operator==(const GraphKey& o) const
    {
      return llr == o.llr && out == o.out && noi == o.noi && ok == o.ok && mi == o.mi && max_iter == o.max_iter && min_iter == o.min_iter &&
             start_iter == o.start_iter && early_stop == o.early_stop && do_extract == o.do_extract;
    }
  };
  struct CachedGraph {
    cudaGraphExec_t exec = nullptr;
    GraphKey        key{};
    uint32_t        launches = 0;
    uint64_t        last_use = 0;
  } graphs[N_GRAPHS];
  uint64_t graph_clock = 0;
  bool      contiguous = false;  // one (K, crc) bucket and code block i at llr offset i*(3K+12), output offset i*K/8
};

// cbs sorted into groups of <= 64 equal (K, crc_kind) blocks
static int build_plan(srsb200_engine* e, uint32_t n, const uint32_t* K, const uint8_t* crc_kind, uint32_t uniform_K, int uniform_kind,
                      const uint64_t* llr_off, const uint64_t* out_off, srsb200_plan** out)
{
  std::map<std::pair<uint32_t, uint32_t>, std::vector<int32_t>> buckets;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t k    = K ? K[i] : uniform_K;
    uint32_t kind = crc_kind ? crc_kind[i] : (uint32_t)uniform_kind;
    if (cbindex_exact(k) < 0) return fail(SRSB200_ERROR_INVALID_INPUTS, "code block %u: K=%u is not an LTE turbo block size", i, k);
    if (kind > 2) return fail(SRSB200_ERROR_INVALID_INPUTS, "code block %u: bad crc kind %u", i, kind);
    buckets[{k, kind}].push_back((int32_t)i);
  }
  srsb200_plan* p = new srsb200_plan();
  p->e            = e;
  p->lane         = e->next_lane;
  e->next_lane    = (e->next_lane + 1) % e->n_lanes;
  p->n_cb         = n;
  uint64_t off    = 0;
  for (auto& kv : buckets) {
    const uint32_t k = kv.first.first, kind = kv.first.second;
    const int kidx = cbindex_exact(k);
    if (ensure_ktable(e, kidx)) {
      delete p;
      return SRSB200_ERROR;
    }
    const uint32_t R = ((k + 3 + W - 1) / W) * W;
    p->max_R = std::max(p->max_R, R);
    const std::vector<int32_t>& ids = kv.second;
    for (size_t s = 0; s < ids.size(); s += 64) {
      Group g;
      g.K = k; g.R = R; g.kidx = (uint32_t)kidx; g.crc_kind = kind; g.ws_off = off; g.wpj = WPJ; g.pad_ = 0;
      size_t cnt = std::min<size_t>(64, ids.size() - s);
      // fill low halves first so that a half-empty group still uses all lanes
      for (int j = 0; j < 64; j++) g.cb[j] = -1;
      size_t nlo = std::min<size_t>(32, cnt);
      // spread: first ceil(cnt/2) blocks in the low halves, the rest in the high halves of the same lanes
      size_t half = (cnt + 1) / 2;
      (void)nlo;
      for (size_t j = 0; j < cnt; j++) {
        if (j < half) g.cb[j] = ids[s + j];
        else g.cb[32 + (j - half)] = ids[s + j];
      }
      off += ((group_ws_words(R) * 4 + 255) / 256) * 256;
      p->h_groups.push_back(g);
    }
  }
  p->n_groups = (uint32_t)p->h_groups.size();
  // 128+ groups fill the machine with job warps even at 16 windows per warp; below that shorter runs cut the latency of the
  // job kernel (a warp walks its run sequentially: 256 steps x ~144 cycles at 16 windows)
  p->wpj = p->n_groups >= 128 ? 16u : (uint32_t)WPJ;
  for (auto& g : p->h_groups) g.wpj = p->wpj;
  // spare slots for regrouped survivors: a quarter of the groups (+ one per range: launch_plan gives range [g0, g1) the slots
  // n_groups + g0 / 4 + s ... n_groups + g1 / 4 + s)
  static const bool no_regroup = getenv("SRSB200_NO_REGROUP") != nullptr;
  if (!no_regroup && buckets.size() == 1 && p->n_groups >= 64) {
    p->n_new = p->n_groups / 4 + srsb200_engine::MAX_SUB + 1;
    Group g  = p->h_groups[0];
    g.wpj    = WPJ;  // few groups after a regrouping: short job-warp runs (latency)
    for (int j = 0; j < 64; j++) g.cb[j] = -1;
    for (uint32_t i = 0; i < p->n_new; i++) {
      g.ws_off = off;
      off += ((group_ws_words(g.R) * 4 + 255) / 256) * 256;
      p->h_groups.push_back(g);
    }
  }
  p->ws_bytes = off;
  p->contiguous = buckets.size() == 1;
  for (uint32_t i = 0; i < n && p->contiguous; i++) {
    const uint32_t k = K ? K[i] : uniform_K;
    if ((llr_off && llr_off[i] != (uint64_t)i * (3ull * k + 12)) || (out_off && out_off[i] != (uint64_t)i * (k / 8))) p->contiguous = false;
  }
  cudaError_t ce;
  const uint32_t n_slots = std::max<uint32_t>(1, p->n_groups + p->n_new);
  if ((ce = cudaMalloc(&p->d_groups, sizeof(Group) * n_slots)) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_ws, std::max<uint64_t>(256, p->ws_bytes))) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_llr_off, sizeof(uint64_t) * std::max<uint32_t>(1, n))) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_out_off, sizeof(uint64_t) * std::max<uint32_t>(1, n))) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_crc_acc, sizeof(uint32_t) * std::max<uint32_t>(1, n))) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_done, std::max<uint32_t>(1, n))) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_active, n_slots)) != cudaSuccess || (ce = cudaMemset(p->d_active, 0, n_slots)) != cudaSuccess ||
      (ce = cudaMalloc(&p->d_arrivals, sizeof(uint32_t) * n_slots)) != cudaSuccess ||
      (ce = cudaMemset(p->d_arrivals, 0, sizeof(uint32_t) * n_slots)) != cudaSuccess ||
      (p->n_new && ((ce = cudaMalloc(&p->d_home, sizeof(int32_t) * std::max<uint32_t>(1, n))) != cudaSuccess ||
                    (ce = cudaMalloc(&p->d_src, sizeof(int32_t) * 64 * p->n_new)) != cudaSuccess ||
                    (ce = cudaMalloc(&p->d_rg_state, sizeof(uint32_t) * srsb200_engine::MAX_SUB)) != cudaSuccess ||
                    (ce = cudaMemset(p->d_rg_state, 0, sizeof(uint32_t) * srsb200_engine::MAX_SUB)) != cudaSuccess))) {
    cudaGetLastError();
    srsb200_plan_destroy(p);
    return fail(SRSB200_ERROR, "plan allocation failed: %s", cudaGetErrorString(ce));
  }
  std::vector<uint64_t> lo(n), oo(n);
  for (uint32_t i = 0; i < n; i++) {
    uint32_t k = K ? K[i] : uniform_K;
    lo[i]      = llr_off ? llr_off[i] : (uint64_t)i * (3ull * k + 12);
    oo[i]      = out_off ? out_off[i] : (uint64_t)i * (k / 8);
  }
  if (n) {
    cudaMemcpyAsync(p->d_groups, p->h_groups.data(), sizeof(Group) * p->h_groups.size(), cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(p->d_llr_off, lo.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream);
    cudaMemcpyAsync(p->d_out_off, oo.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream);
  }
  ce = cudaStreamSynchronize(e->stream);
  if (ce != cudaSuccess) {
    srsb200_plan_destroy(p);
    return fail(SRSB200_ERROR, "plan upload failed: %s", cudaGetErrorString(ce));
  }
  *out = p;
  return SRSB200_SUCCESS;
}

extern "C" void srsb200_plan_destroy(srsb200_plan_t* p)
{
  if (!p) return;
  if (p->e) cudaSetDevice(p->e->device);
  for (auto& cg : p->graphs)
    if (cg.exec) cudaGraphExecDestroy(cg.exec);
  cudaFree(p->d_groups);
  cudaFree(p->d_ws);
  cudaFree(p->d_llr_off);
  cudaFree(p->d_out_off);
  cudaFree(p->d_out_len);
  cudaFree(p->d_cb_max_iter);
  cudaFree(p->d_crc_acc);
  cudaFree(p->d_done);
  cudaFree(p->d_active);
  cudaFree(p->d_arrivals);
  cudaFree(p->d_home);
  cudaFree(p->d_src);
  cudaFree(p->d_rg_state);
  delete p;
}

extern "C" int srsb200_plan_regroup_points(srsb200_engine_t* e, srsb200_plan_t* plan, uint32_t* points, uint32_t n)
{
  if (!e || (!points && n)) return fail(SRSB200_ERROR_INVALID_INPUTS, "null argument");
  for (uint32_t i = 0; i < n; i++) points[i] = 0;
  if (!plan) plan = e->tb_plan;  // NULL: the plan of the engine's latest transport-block submission
  if (!plan || !plan->n_new) return 0;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  uint32_t st[srsb200_engine::MAX_SUB];
  CUDA_TRY(cudaMemcpy(st, plan->d_rg_state, sizeof(st), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n && i < (uint32_t)srsb200_engine::MAX_SUB; i++) points[i] = st[i];
  return srsb200_engine::MAX_SUB;
}

extern "C" int srsb200_tdec_plan_uniform(srsb200_engine_t* e, uint32_t n, uint32_t K, int crc_kind, srsb200_plan_t** plan)
{
  if (!e || !plan) return fail(SRSB200_ERROR_INVALID_INPUTS, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  int r = build_plan(e, n, nullptr, nullptr, K, crc_kind, nullptr, nullptr, plan);
  if (r == SRSB200_SUCCESS) (*plan)->uniform = true;
  return r;
}

struct RangeArgs {
  uint32_t       g0, g1;
  cudaStream_t   st;
  // regrouping (plans with spare slots, device-resident full decodes): the range's slots [slot0, slot0 + cap) of the group table,
  // its state word, and whether a regrouping point has been enqueued yet (from then on the launches cover the slots too)
  uint32_t       slot0 = 0, cap = 0, rg_idx = 0;
  bool           rg = false, rg_started = false;
};

// one launch of the decode chain of a group range; kind: 0 extract, 1 scan, 2 job (+ per-block verdict), 4 emit
// make the engine stream wait for every lazily joined device-resident submission (stream-ordered, does not block the host)
static int join_pending(srsb200_engine* e)
{
  for (int s = 0; s < srsb200_engine::MAX_SUB; s++) {
    if (!e->pending[s]) continue;
    e->pending[s] = false;
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join[s], 0));
  }
  return 0;
}

static void launch_one(srsb200_engine* e, srsb200_plan* p, const RangeArgs& r, int kind, uint32_t n, const int16_t* d_llr, uint32_t max_iter,
                       uint32_t min_iter, int early_stop, uint8_t* d_out, uint8_t* d_noi, uint8_t* d_ok)
{
  const uint32_t ng0 = r.g1 - r.g0;                       // the range's own groups
  const uint32_t ng  = ng0 + (r.rg_started ? r.cap : 0);  // + its slots for regrouped survivors, once a regrouping point has passed
  const uint32_t so  = r.slot0 - r.g0;                    // the slots' index relative to the range's first group
  const Group*   dg = p->d_groups + r.g0;
  uint8_t*       da = p->d_active + r.g0;
  cudaStream_t   st = r.st;
  const int      mode = (n == 0) ? 0 : ((n & 1u) ? 2 : 1);
  const uint32_t nwin_max = (p->max_R + WC - 1) / WC;
  const uint32_t wmin = r.rg_started ? std::min<uint32_t>(p->wpj, WPJ) : p->wpj;  // regrouped groups run WPJ windows per warp
  const dim3     sgrid((ng + 1) / 2), jgrid((nwin_max + 4 * wmin - 1) / (4 * wmin), ng);
  const size_t   jsm = 4 * sizeof(JobWarpSmem);
  int32_t*       home = r.rg ? p->d_home : nullptr;
  uint32_t*      rgs  = r.rg ? p->d_rg_state + r.rg_idx : nullptr;
  switch (kind) {
    case 0: {
      ProfScope ps(e, 0, st);
      extract_kernel<<<dim3(p->max_R / XT, ng0), 256, 0, st>>>(dg, p->d_ws, d_llr, p->d_llr_off, da, p->d_done, p->d_crc_acc, home, rgs, 0u);
    } break;
    case 1: {
      ProfScope ps(e, 5, st);
      if (mode == 0) scan_kernel<0><<<sgrid, 160, sizeof(ScanSmemT<0>), st>>>(dg, p->d_ws, da, ng, ng0, so);
      else if (mode == 1) scan_kernel<1><<<sgrid, 160, sizeof(ScanSmemT<1>), st>>>(dg, p->d_ws, da, ng, ng0, so);
      else scan_kernel<2><<<sgrid, 160, sizeof(ScanSmemT<2>), st>>>(dg, p->d_ws, da, ng, ng0, so);
    } break;
    case 2: {
      ProfScope ps(e, 6, st);
#define SRSB200_JOB(M)                                                                                                                    \
  job_kernel<M><<<jgrid, 128, jsm, st>>>(dg, e->d_ktab, p->d_ws, da, p->d_done, p->d_crc_acc, p->d_arrivals + r.g0, d_noi, d_ok, n + 1, \
                                         max_iter, min_iter, early_stop, p->use_cb_max_iter ? p->d_cb_max_iter : nullptr, ng0, so)
      if (mode == 0) SRSB200_JOB(0); else if (mode == 1) SRSB200_JOB(1); else SRSB200_JOB(2);
#undef SRSB200_JOB
    } break;
    case 3: {
      // regrouping point after half-iteration n (0-based): plan, then - only if the plan kernel regrouped - the channel streams
      // and the state stream of the new groups. n even: the last half-iteration was a DEC1, its output app2 carries the state
      ProfScope ps(e, 10, st);
      const uint32_t attempt = n + 1;
      regroup_plan_kernel<<<1, 1024, 0, st>>>(p->d_groups + r.g0, ng0, so, r.cap, da, p->d_done, p->d_home, p->d_src + 64ull * (r.slot0 - p->n_groups), rgs, attempt);
      extract_kernel<<<dim3(p->max_R / XT, r.cap), 256, 0, st>>>(p->d_groups + r.slot0, p->d_ws, d_llr, p->d_llr_off, p->d_active + r.slot0, p->d_done,
                                                                 p->d_crc_acc, nullptr, rgs, attempt);
      regroup_fill_kernel<<<dim3(p->max_R / 64, r.cap), 256, 0, st>>>(p->d_groups + r.slot0, p->d_active + r.slot0, dg, p->d_src + 64ull * (r.slot0 - p->n_groups),
                                                                      e->d_ktab, p->d_ws, rgs, attempt, (n & 1u) ? 0 : 1);
      e->launches += 2;
    } break;
    default: {
      ProfScope ps(e, 2, st);
      emit_kernel<<<dim3(ng, EMIT_SPLIT), 256, emit_smem_bytes(p->max_R, p->max_R), st>>>(dg, e->d_ktab, p->d_ws, d_noi, d_out, p->d_out_off, p->d_out_len, home,
                                                                                          ng0, so);
    } break;
  }
  e->launches++;
}

/*
 * A decode is a chain of dependent launches: extract, then per half-iteration {scan, job}, then emit. Every
 * half-iteration up to max_iter is enqueued; groups whose code blocks are all done exit at once.
 *
 * Host-pointer submissions of a contiguous equal-size batch are cut into S ranges of groups, each with its own stream:
 * H2D of range s+1 overlaps the decode of range s, and its D2H overlaps the decode of range s+2 (the PCIe copies are
 * the long pole of the end-to-end path: 6.1 bytes per information bit). Device-resident submissions run as one chain:
 * overlapping the latency-bound scans of one range with the window jobs of another was measured and does not pay
 * (scan warps sharing a sub-partition with job warps slow down as much as the overlap gains; DESIGN.md section 5 design history,
 * profiles/r02_trailing_experiment.md).
 */
struct HostIO {
  const int16_t* h_llr = nullptr;
  uint8_t *      h_out = nullptr, *h_noi = nullptr, *h_ok = nullptr;
  uint32_t       L = 0, KB = 0;  // int16 per code block in, bytes per code block out
};

static int launch_plan(srsb200_engine* e, srsb200_plan* p, const int16_t* d_llr, uint32_t max_iter, uint32_t min_iter, int early_stop,
                       uint32_t start_iter, bool do_extract, uint8_t* d_out, uint8_t* d_noi, uint8_t* d_ok, const HostIO* io = nullptr, bool lazy = false)
{
  if (p->n_groups == 0) return SRSB200_SUCCESS;
  if (max_iter == 0) max_iter = 1;  // run_all is a do-while (turbodecoder.c:542-546)
  // (extract_kernel re-arms the per-block CRC / done flags and the group's active flag: no memsets here, so that a
  //  lazily joined submission never touches state of a plan that is still running on the other lane)
  uint32_t S = 1;
  if (!e->profiling) S = std::max(1u, std::min((uint32_t)(io ? e->n_sub : e->n_sub_dev), p->n_groups / 16u));
  if (e->profiling) lazy = false;
  const uint32_t sbase = lazy ? (uint32_t)p->lane * (uint32_t)(srsb200_engine::MAX_SUB / e->n_lanes) : 0u;
  if (lazy) S = std::min(S, (uint32_t)(srsb200_engine::MAX_SUB / e->n_lanes));
  RangeArgs rg[srsb200_engine::MAX_SUB];
  // Host-pointer submissions end with the decode of the LAST range after the last copy has landed, and a decode has a
  // latency floor of ~1 ms however small it is (sequential recursions) - so the ranges shrink geometrically towards the
  // end (weights ... 8 8 4 2 1): the tail is one small decode instead of 1/S of the batch.
  uint32_t wsum = 0, wacc = 0, wgt[srsb200_engine::MAX_SUB];
  for (uint32_t s = 0; s < S; s++) {
    wgt[s] = io ? std::min(8u, 1u << (S - 1 - s)) : 1u;
    wsum += wgt[s];
  }
  for (uint32_t s = 0; s < S; s++) {
    rg[s].g0 = (uint32_t)((uint64_t)p->n_groups * wacc / wsum);
    wacc += wgt[s];
    rg[s].g1 = (uint32_t)((uint64_t)p->n_groups * wacc / wsum);
    rg[s].st = (S == 1 && !lazy) ? e->stream : e->sub[sbase + s];
    // regrouping of the survivors: full device-resident decodes of plans that carry spare slots
    if (p->n_new && !io && do_extract && start_iter == 0 && max_iter >= 6) {
      rg[s].rg     = true;
      rg[s].slot0  = p->n_groups + rg[s].g0 / 4 + s;
      rg[s].cap    = rg[s].g1 / 4 - rg[s].g0 / 4 + 1;
      rg[s].rg_idx = (S == 1 && !lazy) ? 0u : sbase + s;
    }
  }
  // Small decodes (one range on the engine stream: a transport block, a single code block) are launch-bound: extract +
  // 2 x max_iter + emit kernels of tens of microseconds each, and in a multi-threaded PHY every launch takes the
  // process-wide driver lock. Their chain is captured once per (plan, arguments) into a CUDA graph and replayed.
  static const bool use_graphs = getenv("SRSB200_NO_GRAPHS") == nullptr;
  if (use_graphs && S == 1 && !lazy && !io && !e->profiling && start_iter == 0 && do_extract) {
    const srsb200_plan::GraphKey key{d_llr, d_out, d_noi, d_ok, p->use_cb_max_iter ? p->d_cb_max_iter : nullptr, max_iter, min_iter, start_iter, early_stop, do_extract ? 1 : 0};
    srsb200_plan::CachedGraph* slot = nullptr;
    for (auto& cg : p->graphs)
      if (cg.exec && cg.key == key) slot = &cg;
    if (!slot) {
      // least recently used slot
      slot = &p->graphs[0];
      for (auto& cg : p->graphs)
        if (!cg.exec) { slot = &cg; break; } else if (cg.last_use < slot->last_use) slot = &cg;
      if (slot->exec) cudaGraphExecDestroy(slot->exec);
      slot->exec = nullptr;
      cudaGraph_t g = nullptr;
      const uint64_t l0 = e->launches;
      if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        if (do_extract) launch_one(e, p, rg[0], 0, 0, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
        for (uint32_t n = start_iter; n < max_iter; n++) {
          launch_one(e, p, rg[0], 1, n, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
          launch_one(e, p, rg[0], 2, n, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
        }
        launch_one(e, p, rg[0], 4, 0, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
        if (cudaStreamEndCapture(e->stream, &g) == cudaSuccess && g && cudaGraphInstantiate(&slot->exec, g, 0) == cudaSuccess) {
          slot->key      = key;
          slot->launches = (uint32_t)(e->launches - l0);
        } else {
          slot->exec = nullptr;
        }
        if (g) cudaGraphDestroy(g);
      }
      e->launches = l0;
      cudaGetLastError();
    }
    if (slot->exec) {
      slot->last_use = ++p->graph_clock;
      CUDA_TRY(cudaGraphLaunch(slot->exec, e->stream));
      e->launches += slot->launches;
      return SRSB200_SUCCESS;
    }
  }
  const auto t_host0 = std::chrono::steady_clock::now();
  static const bool trace_env = getenv("SRSB200_TRACE") != nullptr;
  const bool  trace = trace_env && io && S > 1;
  cudaEvent_t tev[2 * srsb200_engine::MAX_SUB], tev0 = nullptr;
  if (trace) { cudaEventCreate(&tev0); cudaEventRecord(tev0, e->stream); }
  if (S > 1 || lazy) {
    CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
    for (uint32_t s = 0; s < S; s++) CUDA_TRY(cudaStreamWaitEvent(e->sub[sbase + s], e->ev_fork, 0));
  }
  for (uint32_t s = 0; s < S; s++) {
    // group g of a contiguous plan holds code blocks [64 g, 64 g + 64)
    const uint64_t c0 = 64ull * rg[s].g0, c1 = std::min<uint64_t>(64ull * rg[s].g1, p->n_cb);
    if (io) {
      CUDA_TRY(cudaMemcpyAsync(const_cast<int16_t*>(d_llr) + c0 * io->L, io->h_llr + c0 * io->L, (c1 - c0) * io->L * sizeof(int16_t),
                               cudaMemcpyHostToDevice, rg[s].st));
      if (trace) { cudaEventCreate(&tev[2 * s]); cudaEventCreate(&tev[2 * s + 1]); cudaEventRecord(tev[2 * s], rg[s].st); }
    }
    if (do_extract) launch_one(e, p, rg[s], 0, 0, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
    for (uint32_t n = start_iter; n < max_iter; n++) {
      launch_one(e, p, rg[s], 1, n, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
      launch_one(e, p, rg[s], 2, n, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
      // regrouping points after the 4th, 5th, ... half-iteration, as long as at least three more can follow (the first point that
      // finds few enough survivors regroups; a later point would pay the gather for one or two half-iterations)
      static const uint32_t rg_last = [] { const char* v = getenv("SRSB200_REGROUP_LAST"); return v ? (uint32_t)atoi(v) : 0u; }();
      if (rg[s].rg && n + 1 >= 4 && n + 1 <= (rg_last ? rg_last : std::max(4u, max_iter - 3)) && n + 1 < max_iter) {
        launch_one(e, p, rg[s], 3, n, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
        rg[s].rg_started = true;
      }
    }
    launch_one(e, p, rg[s], 4, 0, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok);
    if (io) {
      CUDA_TRY(cudaMemcpyAsync(io->h_out + c0 * io->KB, d_out + c0 * io->KB, (c1 - c0) * io->KB, cudaMemcpyDeviceToHost, rg[s].st));
      CUDA_TRY(cudaMemcpyAsync(io->h_noi + c0, d_noi + c0, c1 - c0, cudaMemcpyDeviceToHost, rg[s].st));
      CUDA_TRY(cudaMemcpyAsync(io->h_ok + c0, d_ok + c0, c1 - c0, cudaMemcpyDeviceToHost, rg[s].st));
      if (trace) cudaEventRecord(tev[2 * s + 1], rg[s].st);
    }
  }
  if (S > 1 || lazy) {
    for (uint32_t s = 0; s < S; s++) {
      CUDA_TRY(cudaEventRecord(e->ev_join[sbase + s], e->sub[sbase + s]));
      if (lazy) e->pending[sbase + s] = true;
      else CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join[sbase + s], 0));
    }
  }
  if (trace) {
    // opt-in timeline of a host-pointer submission (SRSB200_TRACE=1): when each range's input landed / its results left
    fprintf(stderr, "srsb200 trace: all launches enqueued after %.3f ms of host time\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
    cudaStreamSynchronize(e->stream);
    for (uint32_t s = 0; s < S; s++) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, tev0, tev[2 * s]);
      cudaEventElapsedTime(&b, tev0, tev[2 * s + 1]);
      fprintf(stderr, "srsb200 trace: range %u groups [%u,%u) h2d done %.3f ms, results out %.3f ms\n", s, rg[s].g0, rg[s].g1, a, b);
      cudaEventDestroy(tev[2 * s]);
      cudaEventDestroy(tev[2 * s + 1]);
    }
    cudaEventDestroy(tev0);
  }
  CUDA_TRY(cudaGetLastError());
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_tdec_run_plan_dev(srsb200_engine_t* e, srsb200_plan_t* plan, const int16_t* d_llr, uint32_t max_iter,
                                         uint32_t min_iter, int early_stop, uint8_t* d_out, uint8_t* d_noi, uint8_t* d_crc_ok)
{
  if (!e || !plan || !d_llr || !d_out || !d_noi || !d_crc_ok) return fail(SRSB200_ERROR_INVALID_INPUTS, "null argument");
  if ((uintptr_t)d_llr & 1u) return fail(SRSB200_ERROR_INVALID_INPUTS, "d_llr must be 2-byte aligned");
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  // (profiling runs the submission on the engine stream itself: order it after whatever is still in flight on the lanes)
  if (e->profiling && join_pending(e)) return SRSB200_ERROR;
  return launch_plan(e, plan, d_llr, max_iter, min_iter, early_stop, 0, true, d_out, d_noi, d_crc_ok, nullptr, true);
}

extern "C" int srsb200_tdec_batch(srsb200_engine_t* e, uint32_t n, const uint32_t* K, const uint8_t* crc_kind, const int16_t* llr,
                                  const uint64_t* llr_offset, uint64_t llr_len, uint32_t max_iter, uint32_t min_iter, int early_stop,
                                  uint8_t* out_bytes, const uint64_t* out_offset, uint64_t out_len, uint8_t* noi, uint8_t* crc_ok)
{
  if (!e) return fail(SRSB200_ERROR_NO_DEVICE, "no engine (no CUDA device?)");
  if (n == 0) return SRSB200_SUCCESS;
  if (!K || !llr || !llr_offset || !out_bytes || !out_offset || !noi || !crc_ok) return fail(SRSB200_ERROR_INVALID_INPUTS, "null argument");
  static const bool trace = getenv("SRSB200_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_entry = trace ? now() : 0;
  double t_plan = 0, t_enq = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (cbindex_exact(K[i]) < 0) return fail(SRSB200_ERROR_INVALID_INPUTS, "code block %u: K=%u is not an LTE turbo block size", i, K[i]);
    if (llr_offset[i] + 3ull * K[i] + 12 > llr_len) return fail(SRSB200_ERROR_INVALID_INPUTS, "code block %u: LLR range exceeds llr_len", i);
    if (out_offset[i] + K[i] / 8 > out_len) return fail(SRSB200_ERROR_INVALID_INPUTS, "code block %u: output range exceeds out_len", i);
  }
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  srsb200_plan* p = nullptr;
  int r = 0;
  {
    // same shape as the previous submission -> reuse its plan (group table + workspace)
    bool same = e->cached_plan && e->cached_K.size() == n && memcmp(e->cached_K.data(), K, n * sizeof(uint32_t)) == 0 &&
                memcmp(e->cached_loff.data(), llr_offset, n * sizeof(uint64_t)) == 0 &&
                memcmp(e->cached_ooff.data(), out_offset, n * sizeof(uint64_t)) == 0;
    if (same) {
      for (uint32_t i = 0; i < n && same; i++) same = e->cached_kind[i] == (crc_kind ? crc_kind[i] : (uint8_t)SRSB200_CRC_NONE);
    }
    if (!same) {
      if (e->cached_plan) srsb200_plan_destroy(e->cached_plan);
      e->cached_plan = nullptr;
      r = build_plan(e, n, K, crc_kind, 0, SRSB200_CRC_NONE, llr_offset, out_offset, &p);
      if (r) return r;
      e->cached_plan = p;
      e->cached_K.assign(K, K + n);
      e->cached_kind.resize(n);
      for (uint32_t i = 0; i < n; i++) e->cached_kind[i] = crc_kind ? crc_kind[i] : (uint8_t)SRSB200_CRC_NONE;
      e->cached_loff.assign(llr_offset, llr_offset + n);
      e->cached_ooff.assign(out_offset, out_offset + n);
    }
    p = e->cached_plan;
  }
  void *d_llr, *d_out, *d_noi, *d_ok;
  if (ensure_scratch(e, 0, llr_len * sizeof(int16_t), &d_llr) || ensure_scratch(e, 1, out_len, &d_out) || ensure_scratch(e, 2, n, &d_noi) ||
      ensure_scratch(e, 3, n, &d_ok)) {
    return SRSB200_ERROR;
  }
  cudaError_t ce = cudaSuccess;
  if (trace) t_plan = now();
  if (p->contiguous && llr_len == (uint64_t)n * (3ull * K[0] + 12) && out_len == (uint64_t)n * (K[0] / 8)) {
    // chunked copies overlapped with the decode of the neighbouring ranges
    HostIO io;
    io.h_llr = llr; io.h_out = out_bytes; io.h_noi = noi; io.h_ok = crc_ok;
    io.L = 3 * K[0] + 12; io.KB = K[0] / 8;
    r = launch_plan(e, p, (const int16_t*)d_llr, max_iter, min_iter, early_stop, 0, true, (uint8_t*)d_out, (uint8_t*)d_noi, (uint8_t*)d_ok, &io);
  } else {
    ce = cudaMemcpyAsync(d_llr, llr, llr_len * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess)
      r = launch_plan(e, p, (const int16_t*)d_llr, max_iter, min_iter, early_stop, 0, true, (uint8_t*)d_out, (uint8_t*)d_noi, (uint8_t*)d_ok);
    // only [out_offset[i], +K[i]/8) of every code block is written (the header's contract): gaps in the caller's buffer keep
    // their content and never see scratch bytes of an earlier submission. Neighbouring ranges merge into one transfer.
    if (ce == cudaSuccess && r == 0) {
      std::vector<std::pair<uint64_t, uint64_t>> rg(n);
      for (uint32_t i = 0; i < n; i++) rg[i] = {out_offset[i], out_offset[i] + K[i] / 8};
      std::sort(rg.begin(), rg.end());
      uint64_t a = rg[0].first, b = rg[0].second;
      for (uint32_t i = 1; i <= n && ce == cudaSuccess; i++) {
        if (i < n && rg[i].first <= b) {
          b = std::max(b, rg[i].second);
          continue;
        }
        ce = cudaMemcpyAsync(out_bytes + a, (uint8_t*)d_out + a, b - a, cudaMemcpyDeviceToHost, e->stream);
        if (i < n) { a = rg[i].first; b = rg[i].second; }
      }
    }
    if (ce == cudaSuccess && r == 0) ce = cudaMemcpyAsync(noi, d_noi, n, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess && r == 0) ce = cudaMemcpyAsync(crc_ok, d_ok, n, cudaMemcpyDeviceToHost, e->stream);
  }
  if (trace) t_enq = now();
  if (ce == cudaSuccess && r == 0) ce = cudaStreamSynchronize(e->stream);
  if (trace) fprintf(stderr, "srsb200 trace: host: checks+plan %.3f ms, enqueue %.3f ms, wait %.3f ms\n", t_plan - t_entry, t_enq - t_plan, now() - t_enq);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    return fail(SRSB200_ERROR, "batch decode failed: %s", cudaGetErrorString(ce));
  }
  return r;
}

// ------------------------------------------------------------------ per-object decoder (srsran_tdec_*)
struct srsb200_tdec {
  srsb200_engine* e = nullptr;
  uint32_t max_long_cb = 0;
  int      current_cbidx = -1;
  uint32_t current_long_cb = 0;
  int      n_iter = 0;
  srsb200_plan* plan = nullptr;
  int16_t* d_in = nullptr;
  uint8_t *d_out = nullptr, *d_noi = nullptr, *d_ok = nullptr;
};

extern "C" int srsb200_tdec_init(srsb200_tdec_t** h, srsb200_engine_t* e, uint32_t max_long_cb)
{
  if (!h) return SRSB200_ERROR_INVALID_INPUTS;
  *h = nullptr;
  if (!e) return fail(SRSB200_ERROR_NO_DEVICE, "no engine (no CUDA device?)");
  if (max_long_cb > SRSB200_MAX_K) return fail(SRSB200_ERROR, "max_long_cb %u exceeds %d", max_long_cb, SRSB200_MAX_K);
  srsb200_tdec* t = new srsb200_tdec();
  t->e            = e;
  t->max_long_cb  = max_long_cb;
  std::lock_guard<std::mutex> lk(e->mtx);
  cudaSetDevice(e->device);
  if (cudaMalloc(&t->d_in, (3 * SRSB200_MAX_K + 12) * sizeof(int16_t)) != cudaSuccess || cudaMalloc(&t->d_out, SRSB200_MAX_K / 8) != cudaSuccess ||
      cudaMalloc(&t->d_noi, 4) != cudaSuccess || cudaMalloc(&t->d_ok, 4) != cudaSuccess) {
    cudaGetLastError();
    delete t;
    return fail(SRSB200_ERROR, "tdec allocation failed");
  }
  *h = t;
  return SRSB200_SUCCESS;
}

extern "C" void srsb200_tdec_free(srsb200_tdec_t* t)
{
  if (!t) return;
  cudaSetDevice(t->e->device);
  if (t->plan) srsb200_plan_destroy(t->plan);
  cudaFree(t->d_in);
  cudaFree(t->d_out);
  cudaFree(t->d_noi);
  cudaFree(t->d_ok);
  delete t;
}

extern "C" int srsb200_tdec_new_cb(srsb200_tdec_t* t, uint32_t long_cb)
{
  if (!t) return SRSB200_ERROR;
  if (long_cb > t->max_long_cb) return fail(SRSB200_ERROR, "TDEC was initialized for max_long_cb=%u", t->max_long_cb);
  int idx = cbindex_exact(long_cb);
  if (idx < 0) return fail(SRSB200_ERROR, "Invalid code block size %u", long_cb);
  t->n_iter = 0;
  if (t->current_long_cb != long_cb || !t->plan) {
    std::lock_guard<std::mutex> lk(t->e->mtx);
    cudaSetDevice(t->e->device);
    if (t->plan) srsb200_plan_destroy(t->plan);
    t->plan = nullptr;
    int r = build_plan(t->e, 1, nullptr, nullptr, long_cb, SRSB200_CRC_NONE, nullptr, nullptr, &t->plan);
    if (r) return SRSB200_ERROR;
  }
  t->current_cbidx   = idx;
  t->current_long_cb = long_cb;
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_tdec_get_nof_iterations(srsb200_tdec_t* t) { return t ? t->n_iter : 0; }

extern "C" int srsb200_tdec_iteration(srsb200_tdec_t* t, const int16_t* input, uint8_t* output)
{
  if (!t || !input || !output) return SRSB200_ERROR_INVALID_INPUTS;
  if (t->current_cbidx < 0) return fail(SRSB200_ERROR, "Error CB index not set (call srsb200_tdec_new_cb() first");
  srsb200_engine* e = t->e;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  const uint32_t K = t->current_long_cb;
  bool first = t->n_iter == 0;
  if (first) CUDA_TRY(cudaMemcpyAsync(t->d_in, input, (3 * K + 12) * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
  if (!first) {
    // the previous call marked the block done at its own max_iter; re-arm it for one more half-iteration
    CUDA_TRY(cudaMemsetAsync(t->plan->d_done, 0, t->plan->n_cb, e->stream));
    CUDA_TRY(cudaMemsetAsync(t->plan->d_active, 1, t->plan->n_groups, e->stream));
  }
  int r = launch_plan(e, t->plan, t->d_in, (uint32_t)t->n_iter + 1, 1, 0, (uint32_t)t->n_iter, first, t->d_out, t->d_noi, t->d_ok);
  if (r) return r;
  CUDA_TRY(cudaMemcpyAsync(output, t->d_out, K / 8, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  t->n_iter++;
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_tdec_run_all(srsb200_tdec_t* t, const int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  if (!t || !input || !output) return SRSB200_ERROR_INVALID_INPUTS;
  if (srsb200_tdec_new_cb(t, long_cb)) return SRSB200_ERROR;
  srsb200_engine* e = t->e;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  CUDA_TRY(cudaMemcpyAsync(t->d_in, input, (3 * long_cb + 12) * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
  uint32_t iters = nof_iterations ? nof_iterations : 1;  // do-while: at least one (turbodecoder.c:542-546)
  int r = launch_plan(e, t->plan, t->d_in, iters, 1, 0, 0, true, t->d_out, t->d_noi, t->d_ok);
  if (r) return r;
  CUDA_TRY(cudaMemcpyAsync(output, t->d_out, long_cb / 8, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  t->n_iter = (int)iters;
  return SRSB200_SUCCESS;
}

// hard decision of the latest half-iteration again (the static tdec_decision_byte wrapper, turbodecoder.c:370-378)
extern "C" int srsb200_tdec_get_hard_decision(srsb200_tdec_t* t, uint8_t* output)
{
  if (!t || !output) return SRSB200_ERROR_INVALID_INPUTS;
  if (t->current_cbidx < 0 || t->n_iter == 0) return fail(SRSB200_ERROR, "no half-iteration has run on this decoder yet");
  srsb200_engine* e = t->e;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  CUDA_TRY(cudaMemcpyAsync(output, t->d_out, t->current_long_cb / 8, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SRSB200_SUCCESS;
}

// ------------------------------------------------------------------ rate de-matching
extern "C" int srsb200_rm_turbo_gentables(srsb200_engine_t* e)
{
  // tables are built lazily per (cb_idx, rv) on first use; kept for API symmetry with srsran_rm_turbo_gentables
  return e ? SRSB200_SUCCESS : SRSB200_ERROR_NO_DEVICE;
}

extern "C" int srsb200_rm_turbo_rx_lut(srsb200_engine_t* e, const int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx,
                                       uint32_t rv_idx)
{
  if (rv_idx >= 4 || cb_idx >= LTE_NOF_CB_SIZES) {
    printf("Invalid inputs rv_idx=%d, cb_idx=%d\n", rv_idx, cb_idx);
    return SRSB200_ERROR_INVALID_INPUTS;
  }
  if (!e) return fail(SRSB200_ERROR_NO_DEVICE, "no engine (no CUDA device?)");
  if (!input || !output) return SRSB200_ERROR_INVALID_INPUTS;
  if (in_len == 0) return SRSB200_SUCCESS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  if (ensure_rm_table(e, cb_idx, rv_idx)) return SRSB200_ERROR;
  const uint32_t L = 3 * lte_qpp_params[cb_idx].K + 12;
  void *d_e, *d_buf;
  if (ensure_scratch(e, 4, (size_t)in_len * 2, &d_e) || ensure_scratch(e, 5, (size_t)L * 2, &d_buf)) return SRSB200_ERROR;
  CUDA_TRY(cudaMemcpyAsync(d_e, input, (size_t)in_len * 2, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(d_buf, output, (size_t)L * 2, cudaMemcpyHostToDevice, e->stream));
  RmJob job;
  job.e = (const int16_t*)d_e; job.buf = (int16_t*)d_buf; job.table = e->d_rm_inv[cb_idx][rv_idx]; job.E = in_len; job.L = L;
  job.scramble = 0; job.c_off = 0; job.x1 = 0; job.x2 = 0;
  void* d_job;
  if (ensure_scratch(e, 6, sizeof(RmJob), &d_job)) return SRSB200_ERROR;
  CUDA_TRY(cudaMemcpyAsync(d_job, &job, sizeof(job), cudaMemcpyHostToDevice, e->stream));
  rm_rx_kernel<false><<<1, RM_THREADS, (std::min(in_len, RM_SMEM_ELEMS) + 8) * sizeof(int16_t), e->stream>>>((const RmJob*)d_job, e->d_gold);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(output, d_buf, (size_t)L * 2, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return SRSB200_SUCCESS;
}

// ------------------------------------------------------------------ transport blocks (decode_tb)
// ------------------------------------------------------------------ UL-SCH channel de-interleaver
// geometry + RI bookkeeping of one de-interleave (the q / g / ri_scan pointers are set by the caller)
static void deint_job_fill(DeintJob& j, std::vector<uint32_t>& sc, uint32_t Qm, uint32_t H, uint32_t nsymb, const uint32_t* ri_positions, uint32_t nof_ri)
{
  j.Qm = Qm; j.cols = nsymb; j.rows = H / nsymb;
  const uint32_t ng = H * Qm;
  // RI positions -> sorted, de-duplicated scan-order indices (row, column, bit)
  uint32_t max_ri_pos = 0;
  sc.clear();
  for (uint32_t i = 0; i < nof_ri; i++) {
    const uint32_t p = ri_positions[i];
    if (p >= ng) continue;
    const uint32_t col = p / (j.rows * j.Qm), r2 = p % (j.rows * j.Qm), row = r2 / j.Qm, bit = r2 % j.Qm;
    sc.push_back((row * j.cols + col) * j.Qm + bit);
    max_ri_pos = std::max(max_ri_pos, p);
  }
  std::sort(sc.begin(), sc.end());
  sc.erase(std::unique(sc.begin(), sc.end()), sc.end());
  // first non-RI element in scan order (it has rank 0) and the last position the reference's loop writes to g[0]
  uint32_t s0 = 0;
  while (s0 < sc.size() && sc[s0] == s0) s0++;
  const uint32_t row0 = s0 / (j.cols * j.Qm), rem0 = s0 % (j.cols * j.Qm), col0 = rem0 / j.Qm, bit0 = rem0 % j.Qm;
  const uint32_t p0 = row0 * j.Qm + col0 * j.rows * j.Qm + bit0;
  j.p_star = sc.empty() ? p0 : std::max(p0, max_ri_pos);
  j.nri    = (uint32_t)sc.size();
}

extern "C" int srsb200_ulsch_deinterleave(srsb200_engine_t* e, const int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                                          int16_t* g_bits, const uint32_t* ri_positions, uint32_t nof_ri_bits)
{
  if (!e) return fail(SRSB200_ERROR_NO_DEVICE, "no engine (no CUDA device?)");
  if (!q_bits || !g_bits || Qm == 0 || N_pusch_symbs == 0 || H_prime_total == 0 || H_prime_total % N_pusch_symbs || (nof_ri_bits && !ri_positions))
    return SRSB200_ERROR_INVALID_INPUTS;
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  const size_t ng = (size_t)H_prime_total * Qm;
  DeintJob j;
  std::vector<uint32_t> sc;
  deint_job_fill(j, sc, Qm, H_prime_total, N_pusch_symbs, ri_positions, nof_ri_bits);
  void *d_q, *d_g, *d_dj;
  const size_t dj_bytes = (sizeof(DeintJob) + 15) & ~(size_t)15;
  if (ensure_scratch(e, 12, ng * sizeof(int16_t), &d_q) || ensure_scratch(e, 4, ng * sizeof(int16_t), &d_g) ||
      ensure_scratch(e, 13, dj_bytes + sc.size() * 4 + 16, &d_dj))
    return SRSB200_ERROR;
  j.q       = (const int16_t*)d_q;
  j.g       = (int16_t*)d_g;
  j.ri_scan = reinterpret_cast<const uint32_t*>((uint8_t*)d_dj + dj_bytes);
  Stager stg(e);
  if (stg.reserve(ng * sizeof(int16_t) + dj_bytes + sc.size() * 4 + 4096, ng * sizeof(int16_t) + 4096)) return SRSB200_ERROR;
  CUDA_TRY(stg.h2d(d_q, q_bits, ng * sizeof(int16_t), e->stream));
  CUDA_TRY(stg.h2d(d_dj, &j, sizeof(DeintJob), e->stream));
  if (!sc.empty()) CUDA_TRY(stg.h2d((uint8_t*)d_dj + dj_bytes, sc.data(), sc.size() * 4, e->stream));
  CUDA_TRY(stg.flush(e->stream));
  CUDA_TRY(cudaMemsetAsync(d_g, 0, ng * sizeof(int16_t), e->stream));  // the nof_ri_bits values past the data are left stale by the reference
  {
    ProfScope ps(e, 8);
    const uint32_t tpj = std::max(1u, std::min<uint32_t>(128u, (j.rows + DT_ROWS - 1) / DT_ROWS));
    ulsch_deint_kernel<<<tpj, 256, 0, e->stream>>>((const DeintJob*)d_dj, tpj);
    e->launches++;
  }
  CUDA_TRY(stg.d2h(g_bits, d_g, ng * sizeof(int16_t), e->stream));
  CUDA_TRY(stg.flush(e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  stg.finish();
  return SRSB200_SUCCESS;
}

extern "C" int srsb200_demod_soft_demodulate_s(srsb200_engine_t* e, uint32_t mod, const float* symbols, int16_t* llr, uint32_t nsymbols)
{
  if (!e) return fail(SRSB200_ERROR_NO_DEVICE, "no engine (no CUDA device?)");
  if (mod > 4) {
    fprintf(stderr, "Invalid modulation %d\n", mod);
    return SRSB200_ERROR;
  }
  if (!symbols || !llr) return SRSB200_ERROR_INVALID_INPUTS;
  if (nsymbols == 0) return SRSB200_SUCCESS;
  static const uint32_t bps_of[5] = {1, 2, 4, 6, 8};
  std::lock_guard<std::mutex> lk(e->mtx);
  CUDA_TRY(cudaSetDevice(e->device));
  if (join_pending(e)) return SRSB200_ERROR;
  const size_t nb_in = (size_t)nsymbols * 2 * sizeof(float), nb_out = (size_t)nsymbols * bps_of[mod] * sizeof(int16_t);
  void *d_sym, *d_llr, *d_job;
  if (ensure_scratch(e, 14, nb_in, &d_sym) || ensure_scratch(e, 4, nb_out, &d_llr) || ensure_scratch(e, 15, sizeof(DemodJob), &d_job)) return SRSB200_ERROR;
  DemodJob j;
  j.sym = (const float*)d_sym; j.llr = (int16_t*)d_llr; j.nsym = nsymbols; j.mod = mod;
  Stager stg(e);
  if (stg.reserve(nb_in + sizeof(j) + 4096, nb_out + 4096)) return SRSB200_ERROR;
  CUDA_TRY(stg.h2d(d_sym, symbols, nb_in, e->stream));
  CUDA_TRY(stg.h2d(d_job, &j, sizeof(j), e->stream));
  CUDA_TRY(stg.flush(e->stream));
  {
    ProfScope ps(e, 9);
    demod_kernel<<<dim3(std::max(1u, std::min(1024u, (nsymbols + 255) / 256)), 1), 256, 0, e->stream>>>((const DemodJob*)d_job);
    e->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(stg.d2h(llr, d_llr, nb_out, e->stream));
  CUDA_TRY(stg.flush(e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  stg.finish();
  return SRSB200_SUCCESS;
}

#include "win8.inc"
#include "tb_decode.inc"
#include "tb_encode.inc"
#include "multi.inc"
