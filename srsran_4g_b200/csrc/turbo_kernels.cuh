/*
 * turbo_kernels.cuh - sm_100a kernels of the batched LTE turbo decoder.
 *
 * Design (DESIGN.md has the long form):
 *   - bit-exactness against the reference's generic int16 decoder (turbodecoder_gen.c) forbids windowed/warm-up
 *     recursions: every alpha/beta recursion runs exactly and sequentially in k, with 16-bit WRAPPING arithmetic.
 *   - one thread = TWO code blocks packed in the halves of a 32-bit register (VIADD.16x2 / VIMNMX.S16x2 /
 *     VIADDMNMX.S16x2 wrap exactly like the reference's int16 stores - verified by tools/microbench/int16x2_issue.cu);
 *     one warp = a "group" of up to 64 code blocks of equal K in lock-step.
 *   - all soft streams of a group live in HBM as [row k][32 lanes] packed words, so every access of a warp is one
 *     coalesced 128-byte row and the QPP (de)interleaver is a *row* move with a warp-uniform row index.
 *   - a half-iteration (one constituent MAP decode + glue) is two launches:
 *       scan_kernel   : the two inherently sequential recursions, alpha forward and beta backward, one warp each per
 *                       group (fed by a producer warp), storing only checkpoints: beta every CKB steps, alpha every
 *                       WC*WPJ steps. Exact by construction.
 *       job_kernel    : window-parallel. A warp takes WPJ consecutive windows of WC steps, the alpha checkpoint at the
 *                       start and the beta checkpoints inside, recomputes beta into REGISTERS (8 steps at a time),
 *                       runs alpha + LLR + extrinsic/QPP write-back + hard decision + CRC (by linearity) over it.
 *                       Bit-identical because the recursions are deterministic from the checkpointed states. This
 *                       is where ~70% of the instructions are, spread over K/(WC*WPJ) x more warps than groups.
 *                       The last block of a group to finish applies the per-code-block verdict (CRC == 0 from the
 *                       min_iter-th half-iteration on, half-iteration count, early-stop flag); groups whose blocks
 *                       are all done make the later launches exit immediately.
 *     extract_kernel before (de-multiplex the 3K+12 LLRs into the lane-packed streams) and emit_kernel after (hard
 *     bits back to bytes, through the QPP permutation when the last half-iteration was a DEC2).
 *   - input windows are staged HBM -> shared memory with bulk asynchronous copies (cp.async.bulk + mbarrier, SASS
 *     UBLKCP), double-buffered per warp, so warps spend no issue slots on loads.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

constexpr int W        = 64;  // rows per staged chunk of the scan kernels; streams are padded to a multiple of W rows
constexpr int WC       = 16;  // checkpoint spacing = window of one job = one hard-bit word
constexpr int WPJ      = 8;   // smallest number of windows per job warp (sizes the alpha checkpoint array). The number in use
                              // is per plan (Group::wpj): 16 for big batches (swept 8 / 12 / 16 / 32 on the bench workload: 41.6 /
                              // 42.0 / 42.6 / 42.0 Gbit/s), 8 for small ones, where a job warp's sequential run is latency
constexpr int CKB      = 8;   // spacing of the beta checkpoints (one per 8-step register window of the job kernel)
constexpr int LANES    = 32;
constexpr int NEG_INF2 = 0xD8F0D8F0;  // two int16 of -10000 (turbodecoder_gen.c:37)

// ---------------------------------------------------------------- device-side descriptors
struct KTable {
  // per code-block size, padded to R rows:
  const uint32_t* row1;     // DEC1 write-back row per trellis step i: rev[i]
  const uint32_t* row2;     // DEC2 write-back row per trellis step i: fwd[i]
  // CRC by linearity, one 4x16 nibble table per 16-step window: nib[w][j][v] = XOR over the set bits b of v of
  // x^(K-1-pos(16w+4j+b)+24) mod g, pos(i) = i for DEC1 and fwd[i] for DEC2; [crc_kind] (kind 0 aliases kind 2)
  const uint32_t* nib1[3];
  const uint32_t* nib2[3];
  const uint16_t* rev;      // natural position j -> interleaved index i (for the final bit gather)
};

struct Group {
  uint32_t K;
  uint32_t R;          // rows per stream = ceil((K+3)/W)*W
  uint32_t kidx;       // index into the KTable array
  uint32_t crc_kind;   // 0 none, 1 CRC24A, 2 CRC24B
  uint32_t wpj;        // windows per job warp = alpha checkpoint spacing / WC (8 or 16)
  uint32_t pad_;
  uint64_t ws_off;     // byte offset of this group's workspace
  int32_t  cb[64];     // code-block ids: lane l holds cb[l] (low half) and cb[32+l] (high half); -1 = empty
};

// workspace layout of one group (all uint32 [rows][32]):
//   syst[R] par0[R] par1[R] xa1[R] app2[R] systp[R] ckA[(R/(WC*WPJ)+1)*8] ckB[(R/CKB+1)*8] bits1[R/16] bits2[R/16]
__host__ __device__ inline uint64_t group_ws_words(uint32_t R)
{
  return (uint64_t)LANES * (6ull * R + (R / (WC * WPJ) + 1) * 8ull + (R / CKB + 1) * 8ull + 2ull * (R / 16));
}

struct GroupPtrs {
  uint32_t *syst, *par0, *par1, *xa1, *app2, *systp, *ckA, *ckB, *bits1, *bits2;
};
__host__ __device__ inline GroupPtrs group_ptrs(uint8_t* ws, const Group& g)
{
  GroupPtrs p;
  uint32_t* b = reinterpret_cast<uint32_t*>(ws + g.ws_off);
  uint64_t  s = (uint64_t)g.R * LANES;
  p.syst      = b;
  p.par0      = b + s;
  p.par1      = b + 2 * s;
  p.xa1       = b + 3 * s;  // DEC1 input with a-priori, pre-added: syst + app1 (rows >= K: syst, the termination steps)
  p.app2      = b + 4 * s;
  p.systp     = b + 5 * s;  // syst in interleaved order (systp[j] = syst[fwd[j]]), written once by the first DEC1 job
  p.ckA       = b + 6 * s;
  p.ckB       = p.ckA + (uint64_t)(g.R / (WC * WPJ) + 1) * 8 * LANES;
  p.bits1     = p.ckB + (uint64_t)(g.R / CKB + 1) * 8 * LANES;
  p.bits2     = p.bits1 + (uint64_t)(g.R / 16) * LANES;
  return p;
}

// Launches that follow a regrouping point (regroup_plan_kernel below) cover a range's own groups - relative indices
// [0, n_first) - and then the slots of its regrouped survivors, which sit further back in the group table
__device__ __forceinline__ uint32_t seg_index(uint32_t b, uint32_t n_first, uint32_t second_off) { return b < n_first ? b : b - n_first + second_off; }

// ---------------------------------------------------------------- packed int16x2 arithmetic (wrapping)
__device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t psub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
__device__ __forceinline__ uint32_t pmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t pmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
// max(a + b, c) with a wrapping add: one VIADDMNMX.S16x2
__device__ __forceinline__ uint32_t paddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }

__device__ __forceinline__ void normalise(uint32_t (&s)[8])
{
  uint32_t neg = psub(0u, s[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = padd(s[i], neg);
  s[0] = 0u;
}

// backward recursion step (map_gen_beta, turbodecoder_gen.c:71-103): b <- beta[k] (un-normalised).
// Statement order matters for a lone in-order warp: all adds (FMA pipe) first, then the maxes (ALU pipe) whose add
// operand is oldest, so that no instruction issues within 5 cycles of its producer (measured dependent latency 4-5).
__device__ __forceinline__ void beta_step(uint32_t (&b)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = padd(x, y);
  const uint32_t t2 = padd(b[1], x), t3 = padd(b[1], y), t4 = padd(b[2], y), t5 = padd(b[2], x);
  const uint32_t n1 = paddmax(b[0], xy, b[4]);
  const uint32_t n0 = paddmax(b[4], xy, b[0]);
  const uint32_t n6 = paddmax(b[3], xy, b[7]);
  const uint32_t n7 = paddmax(b[7], xy, b[3]);
  const uint32_t n2 = paddmax(b[5], y, t2);
  const uint32_t n3 = paddmax(b[5], x, t3);
  const uint32_t n4 = paddmax(b[6], x, t4);
  const uint32_t n5 = paddmax(b[6], y, t5);
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}

// ---------------------------------------------------------------- bulk async copy + mbarrier (sm_90+/sm_100a PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// forward recursion step without LLR (alpha part of map_gen_alpha, turbodecoder_gen.c:148-191); same ordering rule
__device__ __forceinline__ void alpha_step(uint32_t (&a)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = padd(x, y);
  const uint32_t t1 = padd(a[2], x), t2 = padd(a[5], x), t5 = padd(a[3], x), t6 = padd(a[4], x);
  const uint32_t n0 = paddmax(a[1], xy, a[0]);
  const uint32_t n3 = paddmax(a[6], xy, a[7]);
  const uint32_t n4 = paddmax(a[0], xy, a[1]);
  const uint32_t n7 = paddmax(a[7], xy, a[6]);
  const uint32_t n1 = paddmax(a[3], y, t1);
  const uint32_t n2 = paddmax(a[4], y, t2);
  const uint32_t n5 = paddmax(a[2], y, t5);
  const uint32_t n6 = paddmax(a[5], y, t6);
  a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3; a[4] = n4; a[5] = n5; a[6] = n6; a[7] = n7;
}

// hard decision of both halves: bit 15 / bit 31 set iff the int16 is > 0  (tdec_gen_decision_byte: app > 0)
__device__ __forceinline__ uint32_t positive_mask(uint32_t v)
{
  return padd(pmax(v, 0u), 0x7fff7fffu) & 0x80008000u;
}

// ---------------------------------------------------------------- scan kernel: sequential alpha / beta recursions
#ifdef SCAN_PROBE
__device__ long long g_probe_cycles[1024];
#endif
// staged chunks in flight per scan warp (bulk copies run NS-1 chunks ahead of the recursion), sized so that the four
// warps of a block fit in one SM's shared memory: 2 streams x 3 stages of 8 KB each = 48 KB
template <int MODE> struct ScanCfg {
  static constexpr int NSTR = 2;  // every mode reads x and y: DEC1 with a-priori reads the pre-added stream xa1
  static constexpr int NS   = 3;
};
template <int MODE> struct ScanStageT {
  uint32_t s[ScanCfg<MODE>::NSTR][W][LANES];
};
// shared memory of one scan block: four recursion warps (two groups x {beta, alpha}) + one producer warp
template <int MODE> struct alignas(128) ScanSmemT {
  ScanStageT<MODE> st[4][ScanCfg<MODE>::NS];     // input chunks in flight per recursion warp
  uint32_t         ck[2][2][W / CKB][LANES][8];  // per beta warp: double-buffered checkpoint staging (one bulk flush per chunk)
  uint64_t         full[4][ScanCfg<MODE>::NS];   // producer -> recursion warp: chunk landed (bulk copy complete_tx)
  uint64_t         empty[4][ScanCfg<MODE>::NS];  // recursion warp -> producer: stage may be overwritten
};

__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// scans: stream 0 = x, stream 1 = y in every mode
#define LOAD_XY(st, r)                   \
  uint32_t x = (st).s[0][(r)][lane];     \
  uint32_t y = (st).s[1][(r)][lane];
// jobs: MODE 0 {syst, par0}, MODE 1 {syst, xa1, par0} (x = xa1), MODE 2 {app2, par1, systp}
#define JLOAD_XY(st, r)                                                  \
  uint32_t x = (st).s[(MODE == 1) ? 1 : 0][(r)][lane];                   \
  uint32_t y = (st).s[(MODE == 1) ? 2 : 1][(r)][lane];

// A lone warp issues roughly one instruction every two cycles (tools/microbench/lone_warp_step.cu: 22.7 cycles per
// 13-instruction step), so the scans are written to execute as few instructions per trellis step as possible:
// fully unrolled blocks of N steps whose x / y are fetched from shared memory up front, one checkpoint test per block.
template <int MODE, int N>
__device__ __forceinline__ void beta_block(const ScanStageT<MODE>& st, int r0, uint32_t (&b)[8], uint32_t (*ck)[LANES][8], int lane)
{
  // steps k = chunk base + r0+N-1 .. r0 (rows r0+N-1 .. r0 of the staged chunk); r0 is a multiple of 8
  uint32_t xs[N], ys[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    LOAD_XY(st, r0 + j);
    xs[j] = x;
    ys[j] = y;
  }
#pragma unroll
  for (int j = N - 1; j >= 0; j--) {
    beta_step(b, xs[j], ys[j]);
    if ((j & 3) == 0) {
      // k % 4 == 0 and k < K: checkpoint first (un-normalised, like beta[8*k+i]), then normalise
#ifndef PROBE_NO_CKPT
      if ((j & 7) == 0) {
        // Into shared memory, not global: a global store keeps its source registers busy until it has read them, and
        // normalise() overwrites them right away - a lone in-order warp then stalls ~100 cycles per checkpoint
        // (measured 39.6 vs 27.9 cycles per step, tools/microbench/block_probe.cu). Layout [slot][lane][8].
        uint4* c4 = reinterpret_cast<uint4*>(&ck[(r0 + j) / CKB][lane][0]);
        c4[0]     = make_uint4(b[0], b[1], b[2], b[3]);
        c4[1]     = make_uint4(b[4], b[5], b[6], b[7]);
      }
#endif
      normalise(b);
    }
  }
}

template <int MODE, int N>
__device__ __forceinline__ void alpha_block(const ScanStageT<MODE>& st, int r0, int k0, uint32_t (&a)[8], uint32_t* ckA, int lane, int run_shift)
{
  // steps k = k0+1 .. k0+N (rows r0 .. r0+N-1); k0 is a multiple of 8
  uint32_t xs[N], ys[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    LOAD_XY(st, r0 + j);
    xs[j] = x;
    ys[j] = y;
  }
  // a job warp only needs the alpha state at its first window: one checkpoint per WPJ windows
  if ((k0 & ((1 << run_shift) - 1)) == 0) {  // run = WC * wpj steps (a power of two)
    uint4* ck = reinterpret_cast<uint4*>(ckA + ((size_t)(k0 >> run_shift) * LANES + lane) * 8);
    ck[0]     = make_uint4(a[0], a[1], a[2], a[3]);
    ck[1]     = make_uint4(a[4], a[5], a[6], a[7]);
  }
#pragma unroll
  for (int j = 0; j < N; j++) {
    alpha_step(a, xs[j], ys[j]);
    if ((j & 3) == 3) normalise(a);  // k = k0 + j + 1 is a multiple of 4
  }
}

/*
 * MODE 0: DEC1 on the first half-iteration (no a-priori)   x = syst,          y = par0
 * MODE 1: DEC1 with a-priori                                x = xa1 (= syst + app1, pre-added by the DEC2 job),  y = par0
 * MODE 2: DEC2                                              x = app2,          y = par1
 * grid = ceil(n_groups / 2), block = 160 = five warps: warps 0-3 are the backward (beta) and forward (alpha) recursions of
 * two groups, warp 4 is the PRODUCER that issues every bulk copy (HBM -> shared memory) for them.
 *  - a recursion warp is a lone, issue-limited instruction stream (tools/microbench/lone_warp_step.cu: ~28 cycles per step
 *    for the bare recursion), so everything that is not the recursion is moved off it: the copies are issued by the
 *    producer (full/empty mbarrier pairs per stage), the "has my next chunk landed?" test is issued one chunk early so
 *    its latency hides behind 64 steps of work, checkpoints go to shared memory and leave with one bulk store per chunk.
 *  - four recursion warps per block on purpose: consecutive warps of a block sit on different SM sub-partitions.
 * (Running both recursions in ONE warp is no alternative: a lone warp is issue-limited, not dependency-limited.)
 * Checkpoints: ckB[k/CKB] = un-normalised beta[k] for k = CKB, 2 CKB, ..., K (what alpha step k consumes);
 *              ckA[k/(WC*WPJ)] = alpha state entering step k+1 (post-normalisation) for k = multiples of WC*WPJ.
 */
template <int MODE>
__global__ void __launch_bounds__(160) scan_kernel(const Group* __restrict__ groups, uint8_t* __restrict__ ws, const uint8_t* __restrict__ group_active,
                                                   uint32_t n_groups, uint32_t n_first, uint32_t second_off)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ScanSmemT<MODE>* sm   = reinterpret_cast<ScanSmemT<MODE>*>(smem_raw);
  constexpr int    NS   = ScanCfg<MODE>::NS;
  const int        wid  = threadIdx.x >> 5;
  const int        lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; w++)
      for (int i = 0; i < NS; i++) {
        mbar_init(&sm->full[w][i], 1);
        mbar_init(&sm->empty[w][i], 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (wid == 4) {
    // ---------------- producer: round-robin over the four recursion warps, NS chunks ahead of each
    const uint32_t* src[4][3];
    int             nchunk[4], cK4[4];
    for (int w = 0; w < 4; w++) {
      const uint32_t bi = blockIdx.x * 2 + (w >> 1), gi = seg_index(bi, n_first, second_off);
      nchunk[w] = 0;
      cK4[w]    = 0;
      if (bi < n_groups && group_active[gi]) {
        const Group&    g  = groups[gi];
        const GroupPtrs gp = group_ptrs(ws, g);
        src[w][0] = (MODE == 2) ? gp.app2 : (MODE == 1 ? gp.xa1 : gp.syst);
        src[w][1] = (MODE == 2) ? gp.par1 : gp.par0;
        src[w][2] = nullptr;
        cK4[w]    = (int)g.K / W;
        nchunk[w] = (w & 1) ? ((int)g.K + W - 1) / W : cK4[w] + 1;
      }
    }
    const int maxc = max(max(nchunk[0], nchunk[1]), max(nchunk[2], nchunk[3]));
    constexpr uint32_t BYTES = W * LANES * 4;
    for (int it = 0; it < maxc; it++) {
#pragma unroll
      for (int w = 0; w < 4; w++) {
        if (it >= nchunk[w]) continue;
        const int s = it % NS;
        if (it >= NS) mbar_wait(&sm->empty[w][s], (uint32_t)(it / NS - 1) & 1u);
        const int      chunk = (w & 1) ? it : cK4[w] - it;
        const size_t   off   = (size_t)chunk * W * LANES;
        const uint32_t bar   = smem_u32(&sm->full[w][s]);
        const uint32_t d0    = smem_u32(&sm->st[w][s].s[0][0][0]);
#ifdef PROBE_NO_COPY
        if (lane == 0) mbar_arrive(&sm->full[w][s]);
        continue;
#endif
        {
          asm volatile(
              "{\n.reg .pred p;\n"
              "elect.sync _|p, 0xffffffff;\n"
              "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
              "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%5], %4, [%0];\n"
              "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%6], %4, [%0];\n}"
              ::"r"(bar), "r"(2u * BYTES), "r"(d0), "r"(d0 + BYTES), "r"(BYTES), "l"(src[w][0] + off), "l"(src[w][1] + off)
              : "memory");
        }
      }
    }
    return;
  }

  // ---------------- recursion warps
  const uint32_t bi  = blockIdx.x * 2 + (wid >> 1), gi = seg_index(bi, n_first, second_off);
  const int      dir = wid & 1;  // 0 = backward (beta), 1 = forward (alpha)
  if (bi >= n_groups || !group_active[gi]) return;
#ifdef SCAN_PROBE
  const long long probe_t0 = clock64();
#endif
  const Group&    g  = groups[gi];
  const GroupPtrs gp = group_ptrs(ws, g);
  const int       K  = (int)g.K;
  uint64_t*       full  = sm->full[wid];
  uint64_t*       empty = sm->empty[wid];
  const int       nch   = dir ? (K + W - 1) / W : K / W + 1;
  // chunk it lives in stage it % NS; `ready` = result of the early test for the chunk about to be consumed
  uint32_t ready = 0;
#ifdef SCAN_PROBE
  long long probe_wait = 0;
  int       probe_nwait = 0;
#endif
  auto acquire = [&](int it) -> const ScanStageT<MODE>& {
#ifdef SCAN_PROBE
    if (!ready) {
      long long w0 = clock64();
      mbar_wait(&full[it % NS], (uint32_t)(it / NS) & 1u);
      probe_wait += clock64() - w0;
      probe_nwait++;
    }
#else
    if (!ready) mbar_wait(&full[it % NS], (uint32_t)(it / NS) & 1u);
#endif
    // test the NEXT chunk now; the answer is only looked at after this chunk's 64 steps
    ready = (it + 1 < nch) ? mbar_test(&full[(it + 1) % NS], (uint32_t)((it + 1) / NS) & 1u) : 0u;
    return sm->st[wid][it % NS];
  };
  auto release = [&](int it) {
    __syncwarp();  // every lane is done reading the stage
    if (lane == 0) mbar_arrive(&empty[it % NS]);
  };

  if (dir == 0) {
    // ---------------- backward recursion (map_gen_beta): chunks cK, cK-1, ..., 0
    const int cK = K / W;  // chunk holding rows K..K+2
    uint32_t  b[8];
    b[0] = 0u;
#pragma unroll
    for (int i = 1; i < 8; i++) b[i] = NEG_INF2;
    for (int it = 0; it <= cK; it++) {
      const int               c  = cK - it;
      const ScanStageT<MODE>& st = acquire(it);
      uint32_t(*ck)[LANES][8]    = sm->ck[wid >> 1][it & 1];
      // the flush of two chunks ago (same staging buffer) must have finished reading shared memory
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      int nslot = W / CKB;
      if (c == cK) {
        // termination steps k = K+2, K+1, K (no a-priori there: xa1 rows >= K hold the bare systematic values); no normalisation at k = K
#pragma unroll
        for (int r = 2; r >= 0; r--) {
          LOAD_XY(st, (K - cK * W) + r);
          beta_step(b, x, y);
        }
        nslot     = (K - cK * W) / CKB + 1;
        uint4* c4 = reinterpret_cast<uint4*>(&ck[nslot - 1][lane][0]);
        c4[0]     = make_uint4(b[0], b[1], b[2], b[3]);
        c4[1]     = make_uint4(b[4], b[5], b[6], b[7]);
      }
      int top = min(K, (c + 1) * W);  // trellis steps of this chunk: k in [c*W, top)
      if (top & 8) {
        top -= 8;
        beta_block<MODE, 8>(st, top - c * W, b, ck, lane);
      }
#ifdef SCAN_PROBE
      long long pb0 = clock64();
#endif
#ifdef PROBE_B8
      for (; top > c * W; top -= 8) beta_block<MODE, 8>(st, top - 8 - c * W, b, ck, lane);
#else
      for (; top > c * W; top -= 16) beta_block<MODE, 16>(st, top - 16 - c * W, b, ck, lane);
#endif
#ifdef SCAN_PROBE
      probe_wait += clock64() - pb0;  // (reported in the "waited" column for the beta warp: time inside the 16-step blocks)
#endif
      release(it);
      // flush this chunk's checkpoints (slots 8c .. 8c+nslot-1 are adjacent in ckB) with one bulk store
#ifdef PROBE_NO_CKPT
      continue;
#endif
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;" ::"l"(
                         gp.ckB + (size_t)c * (W / CKB) * LANES * 8),
                     "r"(smem_u32(&ck[0][0][0])), "r"((uint32_t)nslot * LANES * 8 * 4)
                     : "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ---------------- forward recursion (alpha part of map_gen_alpha): chunks 0, 1, ...
    uint32_t  a[8];
    const int run_shift = 31 - __clz(WC * (int)g.wpj);  // log2 of the alpha checkpoint spacing of this plan
    a[0] = 0u;
#pragma unroll
    for (int i = 1; i < 8; i++) a[i] = NEG_INF2;
    for (int c = 0; c < nch; c++) {
      const ScanStageT<MODE>& st  = acquire(c);
      const int               end = min(K, (c + 1) * W);
      int                     k0  = c * W;
      for (; k0 + 16 <= end; k0 += 16) alpha_block<MODE, 16>(st, k0 - c * W, k0, a, gp.ckA, lane, run_shift);
      if (k0 < end) alpha_block<MODE, 8>(st, k0 - c * W, k0, a, gp.ckA, lane, run_shift);
      release(c);
    }
  }
#ifdef SCAN_PROBE
  if (lane == 0) {
    g_probe_cycles[gi * 2 + dir]       = clock64() - probe_t0;
    g_probe_cycles[512 + gi * 2 + dir] = probe_wait;
    g_probe_cycles[768 + gi * 2 + dir] = probe_nwait;
  }
#endif
}

// ---------------------------------------------------------------- job kernel: window-parallel recompute + LLR + glue
struct JobStage {
  uint32_t s[3][WC][LANES];     // input rows of the window
  uint32_t ck[2][LANES][8];     // un-normalised beta at the top of its two 8-step halves (checkpoints 2w+1, 2w+2)
  uint32_t row[WC];             // write-back row per step
  uint32_t nib[4][16];          // CRC nibble table of this window
};
struct alignas(128) JobWarpSmem {
  JobStage st[2];
  uint64_t bar[2];
};

// pure form of the backward step: d = beta[k] from s = beta[k+1] (possibly normalised)
__device__ __forceinline__ void beta_step_to(uint32_t (&d)[8], const uint32_t (&s)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = padd(x, y);
  d[0] = paddmax(s[4], xy, s[0]);
  d[1] = paddmax(s[0], xy, s[4]);
  d[2] = paddmax(s[5], y, padd(s[1], x));
  d[3] = paddmax(s[5], x, padd(s[1], y));
  d[4] = paddmax(s[6], x, padd(s[2], y));
  d[5] = paddmax(s[6], y, padd(s[2], x));
  d[6] = paddmax(s[3], xy, s[7]);
  d[7] = paddmax(s[7], xy, s[3]);
}
__device__ __forceinline__ void normalise_to(uint32_t (&d)[8], const uint32_t (&s)[8])
{
  const uint32_t neg = psub(0u, s[0]);
  d[0] = 0u;
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = padd(s[i], neg);
}

// forward recursion + LLR step (map_gen_alpha, turbodecoder_gen.c:135-194) with beta[k] in registers; returns m1 - m0
__device__ __forceinline__ uint32_t alpha_llr_step(uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t x, uint32_t y)
{
  uint32_t xy = padd(x, y);
  // information bit 0 branches into state i (m_b[i]) and information bit 1 branches (new[i])
  uint32_t z0 = a[0], z1 = padd(a[3], y), z2 = padd(a[4], y), z3 = a[7];
  uint32_t z4 = a[1], z5 = padd(a[2], y), z6 = padd(a[5], y), z7 = a[6];
  uint32_t o0 = padd(a[1], xy), o1 = padd(a[2], x), o2 = padd(a[5], x), o3 = padd(a[6], xy);
  uint32_t o4 = padd(a[0], xy), o5 = padd(a[3], x), o6 = padd(a[4], x), o7 = padd(a[7], xy);
  // The job kernel is bound by the ALU pipe (max-class instructions), the FMA pipe (adds) has room: form the sixteen
  // branch sums with plain adds and reduce them with three-input maxima (VIMNMX3.S16x2): 8 ALU-pipe instructions for
  // the two 8-way maxima instead of 14 with add-max chains.
  const uint32_t p0 = padd(z0, b[0]), p1 = padd(z1, b[1]), p2 = padd(z2, b[2]), p3 = padd(z3, b[3]);
  const uint32_t p4 = padd(z4, b[4]), p5 = padd(z5, b[5]), p6 = padd(z6, b[6]), p7 = padd(z7, b[7]);
  const uint32_t q0 = padd(o0, b[0]), q1 = padd(o1, b[1]), q2 = padd(o2, b[2]), q3 = padd(o3, b[3]);
  const uint32_t q4 = padd(o4, b[4]), q5 = padd(o5, b[5]), q6 = padd(o6, b[6]), q7 = padd(o7, b[7]);
  // (every three-input maximum takes sums only, so ptxas cannot fold an add back into a VIADDMNMX on the ALU pipe)
  const uint32_t m0  = pmax(pmax3(pmax3(p0, p1, p2), p6, p7), pmax3(p3, p4, p5));
  const uint32_t m1  = pmax(pmax3(pmax3(q0, q1, q2), q6, q7), pmax3(q3, q4, q5));
  a[0] = pmax(z0, o0); a[1] = pmax(z1, o1); a[2] = pmax(z2, o2); a[3] = pmax(z3, o3);
  a[4] = pmax(z4, o4); a[5] = pmax(z5, o5); a[6] = pmax(z6, o6); a[7] = pmax(z7, o7);
  return psub(m1, m0);
}

/*
 * grid = (ceil(nwin_max / (4*WPJ)), n_groups), block = 128 (4 warps). Warp j of block bx handles windows
 * [(4*bx + j)*WPJ, +WPJ) of its group. Write-back (turbodecoder_iter.h:104-128 with the vec_sub / vec_lut glue folded in):
 *   MODE 0: app2[rev[i]]  = L            MODE 1: app2[rev[i]] = L - app1[i]           MODE 2: xa1[fwd[i]] = syst[fwd[i]] + L - app2[i]
 * Everything a window needs - its input rows, its two beta checkpoints, its step table - arrives in shared memory through
 * one group of bulk asynchronous copies issued one window ahead; the loop itself performs no global loads.
 */
template <int MODE>
__global__ void __maxnreg__(152) job_kernel(const Group* __restrict__ groups, const KTable* __restrict__ ktabs, uint8_t* __restrict__ ws,
                                            uint8_t* __restrict__ group_active, uint8_t* __restrict__ done, uint32_t* __restrict__ crc_acc,
                                            uint32_t* __restrict__ arrivals, uint8_t* __restrict__ noi, uint8_t* __restrict__ ok, uint32_t cnt,
                                            uint32_t max_iter, uint32_t min_iter, int early_stop, const uint8_t* __restrict__ max_iter_cb,
                                            uint32_t n_first, uint32_t second_off)
{
  const uint32_t gi = seg_index(blockIdx.y, n_first, second_off);
  if (!group_active[gi]) return;
  __shared__ uint32_t s_last;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int       wid  = threadIdx.x >> 5;
  const int       lane = threadIdx.x & 31;
  JobWarpSmem*    sm   = reinterpret_cast<JobWarpSmem*>(smem_raw) + wid;
  const Group&    g    = groups[gi];
  const uint32_t  K    = g.K;
  const int       nwin = (int)((K + WC - 1) / WC);
  const int       wpj  = (int)g.wpj;  // windows per warp of THIS group (regrouped groups run shorter warps: few groups, latency)
  const uint32_t  nblk = (uint32_t)(nwin + 4 * wpj - 1) / (uint32_t)(4 * wpj);  // blocks of the grid that have windows of this group
  if (blockIdx.x >= nblk) return;
  const int       w0   = (int)(blockIdx.x * 4 + wid) * wpj;
  const int       w1   = min(w0 + wpj, nwin);
  const GroupPtrs gp   = group_ptrs(ws, g);
  const KTable&   kt   = ktabs[g.kidx];

  const uint32_t* in0  = (MODE == 2) ? gp.app2 : gp.syst;
  const uint32_t* in1  = (MODE == 2) ? gp.par1 : (MODE == 1 ? gp.xa1 : gp.par0);
  const uint32_t* in2  = (MODE == 2) ? gp.systp : (MODE == 1 ? gp.par0 : nullptr);
  const uint32_t* rowt = (MODE == 2) ? kt.row2 : kt.row1;
  const uint32_t* nibt = (MODE == 2) ? kt.nib2[g.crc_kind] : kt.nib1[g.crc_kind];
  uint32_t*       dst  = (MODE == 2) ? gp.xa1 : gp.app2;
  uint32_t*       bits = (MODE == 2) ? gp.bits2 : gp.bits1;

  if (lane == 0) {
    mbar_init(&sm->bar[0], 1);
    mbar_init(&sm->bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int w) {
    // rows [w*WC, w*WC + WC) of each stream (streams are padded to a multiple of W >= WC rows), beta checkpoints
    // 2w+1 and 2w+2 (adjacent), and the step table: one elected lane, one predicated asm block (warp stays convergent)
    JobStage&          st    = sm->st[w & 1];
    const uint32_t     bar   = smem_u32(&sm->bar[w & 1]);
    const uint32_t     d0    = smem_u32(&st.s[0][0][0]);
    const size_t       off   = (size_t)w * WC * LANES;
    constexpr uint32_t BYTES = WC * LANES * 4, CKBYTES = 2 * 8 * LANES * 4, TABBYTES = WC * 4 + 4 * 16 * 4;
    const uint32_t*    ck    = gp.ckB + (size_t)(2 * w + 1) * 8 * LANES;
    const uint32_t*    rw    = rowt + (size_t)w * WC;
    const uint32_t*    nb    = nibt + (size_t)w * 64;
    if (MODE != 0) {
      asm volatile(
          "{\n.reg .pred p;\n"
          "elect.sync _|p, 0xffffffff;\n"
          "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%5], %3, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%4], [%6], %3, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%7], [%8], %3, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%9], [%10], %11, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%12], [%13], 64, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%14], [%15], 256, [%0];\n}"
          ::"r"(bar), "r"(3u * BYTES + CKBYTES + TABBYTES), "r"(d0), "r"(BYTES), "r"(d0 + BYTES), "l"(in0 + off), "l"(in1 + off),
            "r"(d0 + 2 * BYTES), "l"(in2 + off), "r"(smem_u32(&st.ck[0][0][0])), "l"(ck), "r"(CKBYTES), "r"(smem_u32(&st.row[0])), "l"(rw),
            "r"(smem_u32(&st.nib[0][0])), "l"(nb)
          : "memory");
    } else {
      asm volatile(
          "{\n.reg .pred p;\n"
          "elect.sync _|p, 0xffffffff;\n"
          "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%5], %3, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%4], [%6], %3, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%7], [%8], %9, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%10], [%11], 64, [%0];\n"
          "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%12], [%13], 256, [%0];\n}"
          ::"r"(bar), "r"(2u * BYTES + CKBYTES + TABBYTES), "r"(d0), "r"(BYTES), "r"(d0 + BYTES), "l"(in0 + off), "l"(in1 + off),
            "r"(smem_u32(&st.ck[0][0][0])), "l"(ck), "r"(CKBYTES), "r"(smem_u32(&st.row[0])), "l"(rw), "r"(smem_u32(&st.nib[0][0])), "l"(nb)
          : "memory");
    }
  };
  if (w0 < nwin) issue(w0);

  const int cb_lo = g.cb[lane], cb_hi = g.cb[32 + lane];
  uint32_t  keep  = 0;
  if (cb_lo < 0 || done[cb_lo]) keep |= 0x0000ffffu;
  if (cb_hi < 0 || done[cb_hi]) keep |= 0xffff0000u;
  uint32_t crc_lo = 0, crc_hi = 0;
  uint32_t ph0 = 0, ph1 = 0;
  // alpha state entering the first window (the windows of this warp are consecutive, so it simply carries on)
  uint32_t a[8];
  if (w0 < nwin) {
    const uint4* ca = reinterpret_cast<const uint4*>(gp.ckA + ((size_t)(w0 / wpj) * LANES + lane) * 8);
    const uint4  c0 = ca[0], c1 = ca[1];
    a[0] = c0.x; a[1] = c0.y; a[2] = c0.z; a[3] = c0.w;
    a[4] = c1.x; a[5] = c1.y; a[6] = c1.z; a[7] = c1.w;
  }

  for (int w = w0; w < w1; w++) {
    const int lo = w * WC;
    const int hi = (int)min((uint32_t)lo + WC, K);
    if (w + 1 < w1) {
      __syncwarp();  // every lane has finished reading the stage the next copy overwrites (window w-1)
      issue(w + 1);
    }
    if (w & 1) { mbar_wait(&sm->bar[1], ph1 & 1u); ph1++; } else { mbar_wait(&sm->bar[0], ph0 & 1u); ph0++; }
    const JobStage& st   = sm->st[w & 1];
    const int       nsub = (hi - lo) / 8;
    uint32_t        bitacc = 0;
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
      // recompute beta[k] for the eight steps k = lo+8s+1 .. lo+8s+8 into registers: B[j] = beta[lo+8s+1+j]
      uint32_t B[8][8];
      {
        const uint4* c  = reinterpret_cast<const uint4*>(&st.ck[s][lane][0]);
        const uint4  c0 = c[0], c1 = c[1];
        B[7][0] = c0.x; B[7][1] = c0.y; B[7][2] = c0.z; B[7][3] = c0.w;
        B[7][4] = c1.x; B[7][5] = c1.y; B[7][6] = c1.z; B[7][7] = c1.w;
      }
      {
        uint32_t t[8];
        if ((uint32_t)(lo + 8 * s + 8) < K) {
          normalise_to(t, B[7]);  // the recursion continues from the normalised state (k % 4 == 0 and k < K)
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) t[i] = B[7][i];
        }
        {
          JLOAD_XY(st, 8 * s + 7);
          beta_step_to(B[6], t, x, y);
        }
      }
#pragma unroll
      for (int j = 5; j >= 3; j--) {
        JLOAD_XY(st, 8 * s + 1 + j);
        beta_step_to(B[j], B[j + 1], x, y);
      }
      {
        uint32_t t[8];
        normalise_to(t, B[3]);  // k = lo + 8s + 4
        JLOAD_XY(st, 8 * s + 3);
        beta_step_to(B[2], t, x, y);
      }
#pragma unroll
      for (int j = 1; j >= 0; j--) {
        JLOAD_XY(st, 8 * s + 1 + j);
        beta_step_to(B[j], B[j + 1], x, y);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int r = 8 * s + j;
        JLOAD_XY(st, r);
        uint32_t L = alpha_llr_step(a, B[j], x, y);
        if (j == 3 || j == 7) normalise(a);
        // what the next half-iteration reads (turbodecoder_iter.h:104-128 with the vec_sub / vec_lut glue folded in; all
        // wrapping int16, so re-association is exact):
        //   MODE 0: app2[rev[i]] = L                        and, once, systp[rev[i]] = syst[i]
        //   MODE 1: app2[rev[i]] = L - app1[i] = (L + syst[i]) - xa1[i]
        //   MODE 2: xa1[fwd[i]]  = syst[fwd[i]] + (L - app2[i]) = systp[i] + (L - app2[i])
        const size_t o = (size_t)st.row[r] * LANES + lane;
        if (MODE == 0) {
          dst[o]      = L;
          gp.systp[o] = x;
        } else if (MODE == 1) {
          dst[o] = psub(padd(L, st.s[0][r][lane]), x);
        } else {
          dst[o] = padd(st.s[2][r][lane], psub(L, x));
        }
        bitacc = ((bitacc >> 1) & 0x7fff7fffu) | positive_mask(L);
      }
    }
    if (nsub == 1) bitacc = (bitacc >> 8) & 0x00ff00ffu;  // K = 8 mod 16: the last word holds eight decisions
    // CRC of the 16 decisions of this window by linearity: four nibble look-ups per code block (the per-bit form costs
    // ~5 ALU-pipe instructions per step, and the ALU pipe is what bounds this kernel)
    {
      const uint32_t wl = bitacc & 0xffffu, wh = bitacc >> 16;
      crc_lo ^= st.nib[0][wl & 15u] ^ st.nib[1][(wl >> 4) & 15u] ^ st.nib[2][(wl >> 8) & 15u] ^ st.nib[3][wl >> 12];
      crc_hi ^= st.nib[0][wh & 15u] ^ st.nib[1][(wh >> 4) & 15u] ^ st.nib[2][(wh >> 8) & 15u] ^ st.nib[3][wh >> 12];
    }
    uint32_t* bw = bits + (size_t)w * LANES + lane;
    if (keep) bitacc = (bitacc & ~keep) | (*bw & keep);  // finished code blocks keep their final decisions
    *bw = bitacc;
  }
  if (cb_lo >= 0 && !(keep & 0xffffu) && crc_lo) atomicXor(&crc_acc[cb_lo], crc_lo);
  if (cb_hi >= 0 && !(keep & 0xffff0000u) && crc_hi) atomicXor(&crc_acc[cb_hi], crc_hi);

  // ---------------- per-code-block verdict, by the LAST block of the group to finish (sch.c:426-456):
  // half-iteration count, CRC == 0 accepted from the min_iter-th on, done flags, group activity for the next launches
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&arrivals[gi], 1u) == nblk - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int active = 0;
  if (threadIdx.x < 64) {
    const int cb = g.cb[threadIdx.x];
    if (cb >= 0) {
      if (!done[cb]) {
        const uint32_t okv = (g.crc_kind != 0 && __ldcg(&crc_acc[cb]) == 0u) ? 1u : 0u;
        noi[cb] = (uint8_t)cnt;
        ok[cb]  = (uint8_t)okv;
        // (max_iter_cb: transport blocks of one submission may carry different limits - one srsran_sch_t each)
        if ((early_stop && okv && cnt >= min_iter) || cnt >= (max_iter_cb ? (uint32_t)max_iter_cb[cb] : max_iter)) done[cb] = 1;
        else active = 1;
      }
      crc_acc[cb] = 0u;
    }
  }
  active = __syncthreads_or(active);
  if (threadIdx.x == 0) {
    group_active[gi] = (uint8_t)(active ? 1 : 0);
    arrivals[gi]     = 0u;
  }
}
#undef LOAD_XY

/*
 * De-multiplex natural-order LLRs (tdec_gen_extract_input, turbodecoder_gen.c:238-258) of up to 64 code blocks into the
 * group's packed [row][lane] streams. grid = (R / XT, n_groups), block = 256. HBM-bound: every LLR is read once
 * (coalesced 192-byte runs per code block) and written once (coalesced 128-byte rows); only the three channel streams are
 * written, plus the termination rows of xa1 (xa1 / app2 / systp rows < K are produced by the decoder before they are read).
 * llr_off[cb] = element offset of the code block's 3K+12 int16 in llr.
 */
constexpr int XT = 64;  // rows per tile
__global__ void __launch_bounds__(256)
extract_kernel(const Group* __restrict__ groups, uint8_t* __restrict__ ws, const int16_t* __restrict__ llr, const uint64_t* __restrict__ llr_off,
               uint8_t* __restrict__ active, uint8_t* __restrict__ done, uint32_t* __restrict__ crc_acc, int32_t* __restrict__ home,
               uint32_t* __restrict__ rg_state, uint32_t rg_attempt)
{
  // rg_attempt != 0: the channel streams of the groups that regrouping point rg_attempt has just formed (nothing is re-armed)
  if (rg_attempt && (*rg_state != rg_attempt || !active[blockIdx.y])) return;
  constexpr int ROWW = 3 * XT / 2 + 1;      // words per tile row: 96 int16 + pad (odd => conflict-free column reads)
  __shared__ uint32_t tile[64][ROWW];
  const Group&    g  = groups[blockIdx.y];
  const uint32_t  K  = g.K;
  const uint32_t  k0 = blockIdx.x * XT;  // first row of this tile
  if (blockIdx.x == 0 && !rg_attempt) {
    // a new decode starts here: re-arm the group and its code blocks (stream-ordered before the first scan)
    if (threadIdx.x < 64 && g.cb[threadIdx.x] >= 0) {
      done[g.cb[threadIdx.x]]    = 0;
      crc_acc[g.cb[threadIdx.x]] = 0;
      if (home) home[g.cb[threadIdx.x]] = (int32_t)blockIdx.y;  // the group whose decision arrays will hold the block's bits
    }
    if (threadIdx.x == 0) active[blockIdx.y] = 1;
    if (threadIdx.x == 0 && blockIdx.y == 0 && rg_state) *rg_state = 0u;
  }
  if (k0 >= g.R) return;
  const GroupPtrs gp   = group_ptrs(ws, g);
  const int       tid  = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t  nbody = (k0 < K) ? min((uint32_t)XT, K - k0) : 0u;  // rows of this tile that are trellis steps (multiple of 8)
  // ---- load: 3*nbody int16 = 1.5*nbody words per code block, contiguous; 8-byte loads (a code block starts on an
  //      8-byte boundary whenever its LLR offset is a multiple of 4 elements: 3K+12 is), four code blocks per warp pass
  const uint32_t nwords = 3 * nbody / 2, npairs = nwords / 2;  // nbody % 8 == 0 => nwords % 12 == 0
  if (nbody) {
    const int sub = lane >> 3, l8 = lane & 7;  // 4 code blocks x 8 lanes
    for (int pass = 0; pass < 2; pass++) {
      const int cbl = wid * 8 + pass * 4 + sub;
      const int cb  = g.cb[cbl];
      if (cb < 0) {
        for (uint32_t i = l8; i < nwords; i += 8) tile[cbl][i] = 0u;
      } else {
        const uint64_t off = llr_off[cb] + 3ull * k0;
        if ((reinterpret_cast<uintptr_t>(llr + off) & 7u) == 0) {  // the ADDRESS, not the offset: the base may be a slice
          const uint2* src = reinterpret_cast<const uint2*>(llr + off);
          for (uint32_t i = l8; i < npairs; i += 8) {
            const uint2 v = __ldg(src + i);
            tile[cbl][2 * i]     = v.x;
            tile[cbl][2 * i + 1] = v.y;
          }
        } else {
          const uint16_t* src = reinterpret_cast<const uint16_t*>(llr + off);
          for (uint32_t i = l8; i < nwords; i += 8) tile[cbl][i] = (uint32_t)src[2 * i] | ((uint32_t)src[2 * i + 1] << 16);
        }
      }
    }
  }
  __syncthreads();
  // ---- store: warp w writes rows w, w+8, ...: lane l packs code blocks l (low half) and 32+l (high half)
  const uint16_t* tlo = reinterpret_cast<const uint16_t*>(&tile[lane][0]);
  const uint16_t* thi = reinterpret_cast<const uint16_t*>(&tile[32 + lane][0]);
#pragma unroll
  for (uint32_t rr = 0; rr < XT / 8; rr++) {
    const uint32_t r = wid + 8 * rr;
    if (r < nbody) {
      const uint32_t v0 = (uint32_t)tlo[3 * r] | ((uint32_t)thi[3 * r] << 16);
      const uint32_t v1 = (uint32_t)tlo[3 * r + 1] | ((uint32_t)thi[3 * r + 1] << 16);
      const uint32_t v2 = (uint32_t)tlo[3 * r + 2] | ((uint32_t)thi[3 * r + 2] << 16);
      const size_t   o  = (size_t)(k0 + r) * LANES + lane;
      gp.syst[o] = v0;
      gp.par0[o] = v1;
      gp.par1[o] = v2;
    }
  }
  // ---- rows >= K of this tile: termination values and zero padding (a handful of rows per group)
  for (uint32_t idx = tid; idx < (XT - nbody) * 32; idx += 256) {
    const uint32_t l = idx & 31, k = k0 + nbody + (idx >> 5);
    if (k >= g.R) break;
    uint32_t sv = 0, p0 = 0, p1 = 0, a2 = 0;
    if (k < K + 3) {
      // syst/par0 from the first six tail values, app2/par1 from the last six
      const uint32_t j = k - K;
      uint16_t       v[2][4];
      for (int h = 0; h < 2; h++) {
        const int cb = g.cb[32 * h + l];
        if (cb >= 0) {
          const int16_t* t = llr + llr_off[cb] + 3ull * K;
          v[h][0] = (uint16_t)t[2 * j];
          v[h][1] = (uint16_t)t[2 * j + 1];
          v[h][2] = (uint16_t)t[6 + 2 * j];
          v[h][3] = (uint16_t)t[6 + 2 * j + 1];
        } else {
          v[h][0] = v[h][1] = v[h][2] = v[h][3] = 0;
        }
      }
      sv = v[0][0] | ((uint32_t)v[1][0] << 16);
      p0 = v[0][1] | ((uint32_t)v[1][1] << 16);
      a2 = v[0][2] | ((uint32_t)v[1][2] << 16);
      p1 = v[0][3] | ((uint32_t)v[1][3] << 16);
    }
    const size_t o = (size_t)k * LANES + l;
    gp.syst[o]  = sv;
    gp.par0[o]  = p0;
    gp.par1[o]  = p1;
    gp.xa1[o]   = sv;  // rows >= K keep the bare systematic value (no a-priori on the termination steps); rows < K are
                       // overwritten by the first DEC2 job before any DEC1 reads them
    gp.app2[o]  = a2;  // second encoder's termination systematic values
  }
}

/*
 * Final hard-decision bytes (tdec_gen_decision_byte, MSB first) of every code block of a group: the decisions of the
 * last half-iteration each block ran; after a DEC2 half-iteration they are gathered through the QPP permutation
 * (app1[fwd[i]] = ext2[i], decision on app1). grid = (n_groups, EMIT_SPLIT), block = 256, dynamic smem =
 * emit_smem_bytes(R, K); block y of a group transposes and emits the code blocks held in lanes [4y, 4y+4).
 * The group's two decision arrays are transposed into shared memory ([code block][16-bit piece]), then each warp
 * produces 32 consecutive output words of one code block (coalesced 128-byte stores).
 */
static constexpr int EMIT_SPLIT = 8;  // blocks per group: block y owns the code blocks in lanes [4y, 4y+4) (both halves)
__host__ __device__ inline size_t emit_smem_bytes(uint32_t R, uint32_t K)
{
  const uint32_t P = R / 16 + 2;  // 16-bit pieces per code block, padded so that a row is an odd number of words
  return (size_t)2 * (64 / EMIT_SPLIT) * (P | 2u) * 2 + ((K + 63) & ~63u) * 2;
}
__global__ void __launch_bounds__(256)
emit_kernel(const Group* __restrict__ groups, const KTable* __restrict__ ktabs, uint8_t* __restrict__ ws, const uint8_t* __restrict__ noi,
            uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_len,
            const int32_t* __restrict__ home, uint32_t n_first, uint32_t second_off)
{
  // home (plans that may regroup): a code block is emitted by the group it ran its last half-iteration in
  extern __shared__ __align__(16) uint8_t esm[];
  constexpr uint32_t NL = LANES / EMIT_SPLIT, NC = 2 * NL;  // lanes / code blocks of this block
  const uint32_t  gi   = seg_index(blockIdx.x, n_first, second_off);
  const Group&    g    = groups[gi];
  const GroupPtrs gp   = group_ptrs(ws, g);
  const uint32_t  K    = g.K;
  const uint32_t  NP   = (K + 15) / 16;             // pieces actually used
  const uint32_t  P    = (g.R / 16 + 2) | 2u;       // row pitch in int16; P/2 is odd => conflict-free transposed stores
  const uint32_t  L0   = NL * blockIdx.y;
  uint16_t*       T1   = reinterpret_cast<uint16_t*>(esm);
  uint16_t*       T2   = T1 + NC * P;
  uint16_t*       srev = T2 + NC * P;
  const int       tid  = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  bool any = false;
  for (uint32_t c = 0; c < NC; c++) {
    const int cb = g.cb[(c / NL) * 32 + L0 + (c % NL)];
    any |= cb >= 0 && (!home || home[cb] == (int32_t)gi);
  }
  if (!any) return;
  for (uint32_t idx = tid; idx < NP * NL; idx += 256) {
    const uint32_t p = idx / NL, l = idx % NL;
    const uint32_t w1 = gp.bits1[(size_t)p * LANES + L0 + l], w2 = gp.bits2[(size_t)p * LANES + L0 + l];
    T1[l * P + p]        = (uint16_t)w1;
    T1[(NL + l) * P + p] = (uint16_t)(w1 >> 16);
    T2[l * P + p]        = (uint16_t)w2;
    T2[(NL + l) * P + p] = (uint16_t)(w2 >> 16);
  }
  const uint16_t* rev = ktabs[g.kidx].rev;
  for (uint32_t i = tid; i < K; i += 256) srev[i] = rev[i];
  __syncthreads();
  const uint32_t nwords = (K + 31) / 32, nchunk = (nwords + 31) / 32;
  for (uint32_t task = wid; task < NC * nchunk; task += 8) {
    const uint32_t c = task / nchunk, word = (task % nchunk) * 32 + lane;
    const int      cb = g.cb[(c / NL) * 32 + L0 + (c % NL)];
    if (cb < 0 || word >= nwords || (home && home[cb] != (int32_t)gi)) continue;
    const uint32_t nt = min(32u, K - word * 32);  // K is a multiple of 8 but not always of 32
    uint32_t       v  = 0;
    if (noi[cb] & 1u) {
      // last half-iteration was a DEC1: natural order, two 16-bit pieces
      v = (uint32_t)T1[c * P + 2 * word] | (nt > 16 ? ((uint32_t)T1[c * P + 2 * word + 1] << 16) : 0u);
    } else {
      const uint16_t* t2 = T2 + c * P;
#pragma unroll 8
      for (uint32_t t = 0; t < 32; t++) {
        const uint32_t tt = (t + lane) & 31u;  // rotate so that the lanes of a warp hit different banks
        if (tt < nt) {
          const uint32_t i = srev[word * 32 + tt];
          v |= (((uint32_t)t2[i >> 4] >> (i & 15u)) & 1u) << tt;
        }
      }
    }
    v = __brev(v);  // bit 31 = step 32*word : MSB-first bytes
    uint8_t* o = out + out_off[cb] + 4ull * word;
    // out_len: bytes of this code block the caller wants (a transport block keeps K/8 - 3 bytes of every code block but
    // the last, because the next block's bytes overwrite the 24 CRC bits: sch.c:430 writes at cb_idx * rlen / 8)
    const uint32_t total = out_len ? out_len[cb] : K / 8;
    const uint32_t nb    = (4 * word >= total) ? 0u : min(nt / 8, total - 4 * word);
    if (nb == 4 && ((reinterpret_cast<uintptr_t>(o) & 3u) == 0)) {
      *reinterpret_cast<uint32_t*>(o) = __byte_perm(v, 0, 0x0123);
    } else {
      for (uint32_t b = 0; b < nb; b++) o[b] = (uint8_t)(v >> (24 - 8 * b));
    }
  }
}

/*
 * Regrouping of the unfinished code blocks. The 64 code blocks of a group run in lock-step, so a group costs a full
 * half-iteration as long as ONE of its blocks is unfinished: on the bench workload 87 % of the blocks are done after five
 * half-iterations, yet every group still runs the sixth. At a regrouping point the survivors of a range of groups are packed
 * into fresh groups (64 per group) and the old groups retire:
 *   regroup_plan_kernel   (one block)  counts the survivors; if they fit the slots set aside for the range and at least halve
 *                         the number of running groups, it writes the new groups' code-block lists, where each lane's state
 *                         comes from (src), each survivor's new home, and swaps the activity flags;
 *   extract_kernel        (rg_attempt) refills the channel streams of the new groups from the caller's natural-order LLRs;
 *   regroup_fill_kernel   moves the ONE stream that carries decoder state across a half-iteration boundary - app2 after a DEC1,
 *                         xa1 after a DEC2 (rows < K; everything else is rewritten before it is read) - and rebuilds systp.
 * The decode of a block is the same sequence of operations on the same values in another lane: results are bit-identical.
 * Only plans of one block size regroup, once per decode and range.
 */
__global__ void __launch_bounds__(1024)
regroup_plan_kernel(Group* __restrict__ groups, uint32_t ng, uint32_t second_off, uint32_t cap, uint8_t* __restrict__ active,
                    const uint8_t* __restrict__ done, int32_t* __restrict__ home, int32_t* __restrict__ src, uint32_t* __restrict__ rg_state,
                    uint32_t rg_attempt)
{
  // One block; a thread owns 16 consecutive lanes of a group (a quarter of its code-block list: four 128-bit loads), all loads of
  // a pass are in flight together - the kernel sits alone on the critical path of its range, so its latency is what counts.
  __shared__ uint32_t s_warp[32], s_tot[2];
  if (*rg_state) return;  // regrouped at an earlier point of this decode
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  constexpr uint32_t PER = 16, CHUNK = 1024 * PER;  // list entries per thread / per pass of the block
  // survivors among this thread's 16 entries of pass c0: bit i set = entry i is an unfinished code block; cb[] = the entries
  auto survivors = [&](uint32_t c0, int (&cb)[PER]) -> uint32_t {
    const uint32_t e0 = c0 + tid * PER, g = e0 >> 6;
    if (g >= ng || !active[g]) return 0u;
    const int4* lp = reinterpret_cast<const int4*>(&groups[g].cb[e0 & 63u]);
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const int4 t = lp[v];
      cb[4 * v] = t.x; cb[4 * v + 1] = t.y; cb[4 * v + 2] = t.z; cb[4 * v + 3] = t.w;
    }
    uint32_t d[PER], m = 0;
#pragma unroll
    for (int i = 0; i < (int)PER; i++) d[i] = cb[i] >= 0 ? (uint32_t)done[cb[i]] : 1u;
#pragma unroll
    for (int i = 0; i < (int)PER; i++) m |= (d[i] ? 0u : 1u) << i;
    return m;
  };
  if (tid < 2) s_tot[tid] = 0u;
  __syncthreads();
  {
    uint32_t c = 0, a = 0;
    for (uint32_t c0 = 0; c0 < ng * 64; c0 += CHUNK) {
      int cb[PER];
      c += __popc(survivors(c0, cb));
      const uint32_t g = (c0 + tid * PER) >> 6;
      a += ((tid & 3u) == 0 && g < ng && active[g]) ? 1u : 0u;  // a group is active as long as one of its blocks is unfinished
    }
    c = __reduce_add_sync(0xffffffffu, c);
    a = __reduce_add_sync(0xffffffffu, a);
    if (lane == 0 && c) atomicAdd(&s_tot[0], c);
    if (lane == 0 && a) atomicAdd(&s_tot[1], a);
  }
  __syncthreads();
  const uint32_t total = s_tot[0], running = s_tot[1], ngn = (total + 63) / 64;
  if (total == 0 || ngn > cap || 2 * ngn > running) {
    for (uint32_t i = tid; i < cap; i += 1024) active[second_off + i] = 0;  // slots of an earlier decode stay retired
    return;
  }
  uint32_t carry = 0;
  for (uint32_t c0 = 0; c0 < ng * 64; c0 += CHUNK) {
    int            cb[PER];
    const uint32_t m = survivors(c0, cb), mine = __popc(m);
    // exclusive scan of `mine` over the block: shuffles within a warp, the 32 warp totals through shared memory
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
      if ((int)lane >= d) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t wbase = 0, ctot = 0;
    for (uint32_t w = 0; w < 32; w++) {
      const uint32_t v = s_warp[w];
      if (w < wid) wbase += v;
      ctot += v;
    }
    uint32_t       q  = carry + wbase + inc - mine;
    const uint32_t e0 = c0 + tid * PER, g = e0 >> 6;
#pragma unroll
    for (int i = 0; i < (int)PER; i++) {
      if ((m >> i) & 1u) {
        groups[second_off + (q >> 6)].cb[q & 63u] = cb[i];
        src[q]       = (int32_t)((g << 6) | ((e0 & 63u) + (uint32_t)i));
        home[cb[i]]  = (int32_t)(second_off + (q >> 6));
        q++;
      }
    }
    carry += ctot;
    __syncthreads();  // s_warp is rewritten by the next pass; every thread has read active[] of this pass
    if ((tid & 3u) == 0 && g < ng) active[g] = 0;
  }
  for (uint32_t q = total + tid; q < ngn * 64; q += 1024) {  // the empty lanes of the last new group
    groups[second_off + (q >> 6)].cb[q & 63u] = -1;
    src[q] = -1;
  }
  for (uint32_t i = tid; i < cap; i += 1024) active[second_off + i] = i < ngn ? 1 : 0;
  if (tid == 0) *rg_state = rg_attempt;
}

// grid = (R / 64, slots of the range), block = 256: eight rows x 32 lanes per pass. new_groups / new_active / src start at the
// range's first slot, old_groups at the range's first group; from_app2 = the last half-iteration was a DEC1
__global__ void __launch_bounds__(256)
regroup_fill_kernel(const Group* __restrict__ new_groups, const uint8_t* __restrict__ new_active, const Group* __restrict__ old_groups,
                    const int32_t* __restrict__ src, const KTable* __restrict__ ktabs, uint8_t* __restrict__ ws, const uint32_t* __restrict__ rg_state,
                    uint32_t rg_attempt, int from_app2)
{
  if (*rg_state != rg_attempt || !new_active[blockIdx.y]) return;
  const Group&    g    = new_groups[blockIdx.y];
  const GroupPtrs gp   = group_ptrs(ws, g);
  const uint32_t  K    = g.K, lane = threadIdx.x & 31u, rsub = threadIdx.x >> 5;
  const uint32_t* row2 = ktabs[g.kidx].row2;
  // the two code blocks of this lane: 16-bit column (slot & 31, half slot >> 5) of the state stream of their old group
  const uint16_t* col[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int32_t sv = src[blockIdx.y * 64 + 32 * h + lane];
    col[h]           = nullptr;
    if (sv >= 0) {
      const GroupPtrs op = group_ptrs(ws, old_groups[(uint32_t)sv >> 6]);
      col[h] = reinterpret_cast<const uint16_t*>(from_app2 ? op.app2 : op.xa1) + 2u * ((uint32_t)sv & 31u) + (((uint32_t)sv >> 5) & 1u);
    }
  }
  uint32_t* dst = from_app2 ? gp.app2 : gp.xa1;
  uint32_t  lo[8], hi[8], sp[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t r = blockIdx.x * 64 + rsub + 8 * i;
    lo[i] = hi[i] = sp[i] = 0u;
    if (r < K) {
      if (col[0]) lo[i] = col[0][(size_t)r * 64];
      if (col[1]) hi[i] = col[1][(size_t)r * 64];
      sp[i] = gp.syst[(size_t)row2[r] * LANES + lane];  // systp[j] = syst[fwd[j]]
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t r = blockIdx.x * 64 + rsub + 8 * i;
    if (r < K) {
      dst[(size_t)r * LANES + lane]      = lo[i] | (hi[i] << 16);
      gp.systp[(size_t)r * LANES + lane] = sp[i];
    }
  }
}

}  // namespace srsb200
