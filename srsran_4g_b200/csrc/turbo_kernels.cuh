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
 *   - a half-iteration (one constituent MAP decode + glue) is three launches:
 *       scan_kernel   : the two inherently sequential recursions, alpha forward and beta backward, one warp each per
 *                       group, storing only the state every WC steps (checkpoints). Exact by construction.
 *       job_kernel    : window-parallel. A warp takes a window of WC steps, its alpha checkpoint at the start and its
 *                       beta checkpoint at the end, recomputes beta of the window into REGISTERS (8 steps at a time),
 *                       runs alpha + LLR + extrinsic/QPP write-back + hard decision + CRC (by linearity) over it.
 *                       Bit-identical because the recursions are deterministic from the checkpointed states. This
 *                       is where ~75% of the instructions are, spread over K/WC x more warps than code-block groups.
 *       status_kernel : per-code-block CRC verdict, half-iteration count and early-stop flag; groups whose blocks
 *                       are all done make the later launches exit immediately.
 *   - input windows are staged HBM -> shared memory with bulk asynchronous copies (cp.async.bulk + mbarrier, SASS
 *     UBLKCP), double-buffered per warp, so warps spend no issue slots on loads.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

constexpr int W        = 32;  // rows per staged chunk of the scan kernels; streams are padded to a multiple of W rows
constexpr int WC       = 16;  // checkpoint spacing = window of one job = one hard-bit word
constexpr int WPJ      = 8;   // windows per job warp
constexpr int LANES    = 32;
constexpr int NEG_INF2 = 0xD8F0D8F0;  // two int16 of -10000 (turbodecoder_gen.c:37)

// ---------------------------------------------------------------- device-side descriptors
struct KTable {
  // per code-block size: (row index to scatter to, CRC position word) per trellis step, padded to R rows
  const uint2* dec1[3];  // [crc_kind] : .x = rev[i],  .y = R_kind[K-1-i]
  const uint2* dec2[3];  // [crc_kind] : .x = fwd[i],  .y = R_kind[K-1-fwd[i]]
  const uint16_t* rev;   // natural position j -> interleaved index i (for the final bit gather)
};

struct Group {
  uint32_t K;
  uint32_t R;          // rows per stream = ceil((K+3)/W)*W
  uint32_t kidx;       // index into the KTable array
  uint32_t crc_kind;   // 0 none, 1 CRC24A, 2 CRC24B
  uint64_t ws_off;     // byte offset of this group's workspace
  int32_t  cb[64];     // code-block ids: lane l holds cb[l] (low half) and cb[32+l] (high half); -1 = empty
};

// workspace layout of one group (all uint32 [rows][32]):
//   syst[R] par0[R] par1[R] app1p[R] app2[R] ckA[(R/WC+1)*8] ckB[(R/WC+1)*8] bits1[R/16] bits2[R/16]
__host__ __device__ inline uint64_t group_ws_words(uint32_t R)
{
  return (uint64_t)LANES * (5ull * R + 2ull * (R / WC + 1) * 8ull + 2ull * (R / 16));
}

struct GroupPtrs {
  uint32_t *syst, *par0, *par1, *app1p, *app2, *ckA, *ckB, *bits1, *bits2;
};
__host__ __device__ inline GroupPtrs group_ptrs(uint8_t* ws, const Group& g)
{
  GroupPtrs p;
  uint32_t* b = reinterpret_cast<uint32_t*>(ws + g.ws_off);
  uint64_t  s = (uint64_t)g.R * LANES;
  p.syst      = b;
  p.par0      = b + s;
  p.par1      = b + 2 * s;
  p.app1p     = b + 3 * s;
  p.app2      = b + 4 * s;
  p.ckA       = b + 5 * s;
  p.ckB       = p.ckA + (uint64_t)(g.R / WC + 1) * 8 * LANES;
  p.bits1     = p.ckB + (uint64_t)(g.R / WC + 1) * 8 * LANES;
  p.bits2     = p.bits1 + (uint64_t)(g.R / 16) * LANES;
  return p;
}

// ---------------------------------------------------------------- packed int16x2 arithmetic (wrapping)
__device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t psub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
__device__ __forceinline__ uint32_t pmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
// max(a + b, c) with a wrapping add: one VIADDMNMX.S16x2
__device__ __forceinline__ uint32_t paddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }

__device__ __forceinline__ void normalise(uint32_t (&s)[8])
{
  uint32_t neg = psub(0u, s[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s[i] = padd(s[i], neg);
  s[0] = 0u;
}

// backward recursion step (map_gen_beta, turbodecoder_gen.c:71-103): b <- beta[k] (un-normalised)
__device__ __forceinline__ void beta_step(uint32_t (&b)[8], uint32_t x, uint32_t y)
{
  uint32_t xy = padd(x, y);
  uint32_t n0 = paddmax(b[4], xy, b[0]);
  uint32_t n1 = paddmax(b[0], xy, b[4]);
  uint32_t n2 = paddmax(b[5], y, padd(b[1], x));
  uint32_t n3 = paddmax(b[5], x, padd(b[1], y));
  uint32_t n4 = paddmax(b[6], x, padd(b[2], y));
  uint32_t n5 = paddmax(b[6], y, padd(b[2], x));
  uint32_t n6 = paddmax(b[3], xy, b[7]);
  uint32_t n7 = paddmax(b[7], xy, b[3]);
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}

// forward recursion + LLR step (map_gen_alpha, turbodecoder_gen.c:135-194); returns m1 - m0
__device__ __forceinline__ uint32_t alpha_llr_step(uint32_t (&a)[8], const uint4 bl, const uint4 bh, uint32_t x, uint32_t y)
{
  uint32_t xy = padd(x, y);
  // information bit 0 branches into state i (m_b[i]) and information bit 1 branches (new[i])
  uint32_t z0 = a[0], z1 = padd(a[3], y), z2 = padd(a[4], y), z3 = a[7];
  uint32_t z4 = a[1], z5 = padd(a[2], y), z6 = padd(a[5], y), z7 = a[6];
  uint32_t o0 = padd(a[1], xy), o1 = padd(a[2], x), o2 = padd(a[5], x), o3 = padd(a[6], xy);
  uint32_t o4 = padd(a[0], xy), o5 = padd(a[3], x), o6 = padd(a[4], x), o7 = padd(a[7], xy);
  // two chains per maximum keep the dependent depth short
  uint32_t m0a = padd(z0, bl.x), m0b = padd(z4, bh.x);
  m0a = paddmax(z1, bl.y, m0a); m0b = paddmax(z5, bh.y, m0b);
  m0a = paddmax(z2, bl.z, m0a); m0b = paddmax(z6, bh.z, m0b);
  m0a = paddmax(z3, bl.w, m0a); m0b = paddmax(z7, bh.w, m0b);
  uint32_t m1a = padd(o0, bl.x), m1b = padd(o4, bh.x);
  m1a = paddmax(o1, bl.y, m1a); m1b = paddmax(o5, bh.y, m1b);
  m1a = paddmax(o2, bl.z, m1a); m1b = paddmax(o6, bh.z, m1b);
  m1a = paddmax(o3, bl.w, m1a); m1b = paddmax(o7, bh.w, m1b);
  a[0] = pmax(z0, o0); a[1] = pmax(z1, o1); a[2] = pmax(z2, o2); a[3] = pmax(z3, o3);
  a[4] = pmax(z4, o4); a[5] = pmax(z5, o5); a[6] = pmax(z6, o6); a[7] = pmax(z7, o7);
  return psub(pmax(m1a, m1b), pmax(m0a, m0b));
}

// ---------------------------------------------------------------- bulk async copy + mbarrier (sm_90+/sm_100a PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
  asm volatile("fence.proxy.async;" ::: "memory");
}

// forward recursion step without LLR (alpha part of map_gen_alpha, turbodecoder_gen.c:148-191)
__device__ __forceinline__ void alpha_step(uint32_t (&a)[8], uint32_t x, uint32_t y)
{
  uint32_t xy = padd(x, y);
  uint32_t n0 = paddmax(a[1], xy, a[0]);
  uint32_t n1 = paddmax(a[3], y, padd(a[2], x));
  uint32_t n2 = paddmax(a[4], y, padd(a[5], x));
  uint32_t n3 = paddmax(a[6], xy, a[7]);
  uint32_t n4 = paddmax(a[0], xy, a[1]);
  uint32_t n5 = paddmax(a[2], y, padd(a[3], x));
  uint32_t n6 = paddmax(a[5], y, padd(a[4], x));
  uint32_t n7 = paddmax(a[7], xy, a[6]);
  a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3; a[4] = n4; a[5] = n5; a[6] = n6; a[7] = n7;
}

// hard decision of both halves: bit 15 / bit 31 set iff the int16 is > 0  (tdec_gen_decision_byte: app > 0)
__device__ __forceinline__ uint32_t positive_mask(uint32_t v)
{
  return padd(pmax(v, 0u), 0x7fff7fffu) & 0x80008000u;
}

// ---------------------------------------------------------------- scan kernel: sequential alpha / beta recursions
struct ScanStage {
  uint32_t s[3][W][LANES];
};
struct ScanSmem {
  ScanStage st[2];
  uint64_t  bar[2];
};

struct ScanPipe {
  // double-buffered chunk loader (chunks of W rows); all members are warp-uniform scalars
  ScanSmem*       sm;
  const uint32_t *src0, *src1, *src2;
  int             held0, held1;
  bool            pend0, pend1;
  uint32_t        phase0, phase1;

  __device__ __forceinline__ void issue(int c)
  {
    const int s = c & 1;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      ScanStage& st  = sm->st[s];
      uint64_t*  bar = &sm->bar[s];
      mbar_expect_tx(bar, (src2 ? 3u : 2u) * W * LANES * 4);
      bulk_g2s(&st.s[0][0][0], src0 + (size_t)c * W * LANES, W * LANES * 4, bar);
      bulk_g2s(&st.s[1][0][0], src1 + (size_t)c * W * LANES, W * LANES * 4, bar);
      if (src2) bulk_g2s(&st.s[2][0][0], src2 + (size_t)c * W * LANES, W * LANES * 4, bar);
    }
    if (s) { held1 = c; pend1 = true; } else { held0 = c; pend0 = true; }
  }
  __device__ __forceinline__ void prefetch(int c)
  {
    if (c >= 0 && ((c & 1) ? held1 : held0) != c) issue(c);
  }
  __device__ __forceinline__ const ScanStage& acquire(int c)
  {
    const int s = c & 1;
    if ((s ? held1 : held0) != c) issue(c);
    if (s ? pend1 : pend0) {
      mbar_wait(&sm->bar[s], (s ? phase1 : phase0) & 1u);
      if (s) { phase1++; pend1 = false; } else { phase0++; pend0 = false; }
    }
    return sm->st[s];
  }
};

#define LOAD_XY(st, r)                                                   \
  uint32_t x = (st).s[0][(r)][lane];                                     \
  uint32_t y;                                                            \
  if (MODE == 1) {                                                       \
    x = padd(x, (st).s[1][(r)][lane]);                                   \
    y = (st).s[2][(r)][lane];                                            \
  } else {                                                               \
    y = (st).s[1][(r)][lane];                                            \
  }

/*
 * MODE 0: DEC1 on the first half-iteration (no a-priori)   x = syst,          y = par0
 * MODE 1: DEC1 with a-priori                                x = syst + app1p,  y = par0
 * MODE 2: DEC2                                              x = app2,          y = par1
 * grid = (n_groups, 2): blockIdx.y == 0 runs the backward (beta) recursion, 1 the forward (alpha) recursion.
 * Checkpoints: ckB[ceil(k/WC)] = un-normalised beta[k] for k = multiples of WC and k = K (what alpha step k consumes);
 *              ckA[k/WC] = alpha state entering step k+1 (post-normalisation) for k = multiples of WC.
 */
template <int MODE>
__global__ void __launch_bounds__(32) scan_kernel(const Group* __restrict__ groups, uint8_t* __restrict__ ws, const uint8_t* __restrict__ group_active)
{
  if (!group_active[blockIdx.x]) return;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ScanSmem*       sm   = reinterpret_cast<ScanSmem*>(smem_raw);
  const int       lane = threadIdx.x;
  const Group&    g    = groups[blockIdx.x];
  const GroupPtrs gp   = group_ptrs(ws, g);
  const uint32_t  K    = g.K;
  if (lane == 0) {
    mbar_init(&sm->bar[0], 1);
    mbar_init(&sm->bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  ScanPipe pipe;
  pipe.sm     = sm;
  pipe.phase0 = pipe.phase1 = 0;
  pipe.held0 = pipe.held1 = -1;
  pipe.pend0 = pipe.pend1 = false;
  pipe.src0 = (MODE == 2) ? gp.app2 : gp.syst;
  pipe.src1 = (MODE == 2) ? gp.par1 : (MODE == 1 ? gp.app1p : gp.par0);
  pipe.src2 = (MODE == 1) ? gp.par0 : nullptr;
  const int nwin = (int)((K + WC - 1) / WC);

  if (blockIdx.y == 0) {
    // ---------------- backward recursion (map_gen_beta)
    const int cK = (int)(K / W);  // chunk holding rows K..K+2
    uint32_t  b[8];
    b[0] = 0u;
#pragma unroll
    for (int i = 1; i < 8; i++) b[i] = NEG_INF2;
    pipe.prefetch(cK);
    pipe.prefetch(cK - 1);
    for (int c = cK; c >= 0; c--) {
      const ScanStage& st = pipe.acquire(c);
      if (c == cK) {
        // termination steps k = K+2, K+1, K (no a-priori there: app1p rows >= K stay zero); no normalisation at k = K
#pragma unroll
        for (int r = 2; r >= 0; r--) {
          LOAD_XY(st, (K - (uint32_t)cK * W) + r);
          beta_step(b, x, y);
        }
        uint32_t* ck = gp.ckB + (size_t)nwin * 8 * LANES + lane;
#pragma unroll
        for (int i = 0; i < 8; i++) ck[i * LANES] = b[i];
      }
      const int top = (int)min(K, (uint32_t)(c + 1) * W) - 4;
      for (int kb = top; kb >= c * W; kb -= 4) {
        const int r = kb - c * W;
#pragma unroll
        for (int j = 3; j >= 0; j--) {
          LOAD_XY(st, r + j);
          beta_step(b, x, y);
        }
        // kb % 4 == 0 and kb < K: checkpoint first (un-normalised, like beta[8*k+i]), then normalise
        if ((kb % WC) == 0 && kb > 0) {
          uint32_t* ck = gp.ckB + (size_t)(kb / WC) * 8 * LANES + lane;
#pragma unroll
          for (int i = 0; i < 8; i++) ck[i * LANES] = b[i];
        }
        normalise(b);
      }
      pipe.prefetch(c - 2);
    }
  } else {
    // ---------------- forward recursion (alpha part of map_gen_alpha)
    const int nch = (int)((K + W - 1) / W);
    uint32_t  a[8];
    a[0] = 0u;
#pragma unroll
    for (int i = 1; i < 8; i++) a[i] = NEG_INF2;
    pipe.prefetch(0);
    pipe.prefetch(1 < nch ? 1 : -1);
    for (int c = 0; c < nch; c++) {
      const ScanStage& st  = pipe.acquire(c);
      const int        len = (int)min((uint32_t)W, K - (uint32_t)c * W);
      for (int r = 0; r < len; r += 4) {
        if ((r % WC) == 0) {
          uint32_t* ck = gp.ckA + (size_t)((c * W + r) / WC) * 8 * LANES + lane;
#pragma unroll
          for (int i = 0; i < 8; i++) ck[i * LANES] = a[i];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          LOAD_XY(st, r + j);
          alpha_step(a, x, y);
        }
        normalise(a);  // k = c*W + r + 4 is a multiple of 4
      }
      pipe.prefetch(c + 2 < nch ? c + 2 : -1);
    }
  }
}

// ---------------------------------------------------------------- job kernel: window-parallel recompute + LLR + glue
struct JobStage {
  uint32_t s[3][WC][LANES];
  uint2    tab[WC];
};
struct JobWarpSmem {
  JobStage st[2];
  uint64_t bar[2];
};

// forward recursion + LLR step (map_gen_alpha, turbodecoder_gen.c:135-194) with beta[k] in registers; returns m1 - m0
__device__ __forceinline__ uint32_t alpha_llr_step(uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t x, uint32_t y)
{
  uint32_t xy = padd(x, y);
  // information bit 0 branches into state i (m_b[i]) and information bit 1 branches (new[i])
  uint32_t z0 = a[0], z1 = padd(a[3], y), z2 = padd(a[4], y), z3 = a[7];
  uint32_t z4 = a[1], z5 = padd(a[2], y), z6 = padd(a[5], y), z7 = a[6];
  uint32_t o0 = padd(a[1], xy), o1 = padd(a[2], x), o2 = padd(a[5], x), o3 = padd(a[6], xy);
  uint32_t o4 = padd(a[0], xy), o5 = padd(a[3], x), o6 = padd(a[4], x), o7 = padd(a[7], xy);
  // two chains per maximum keep the dependent depth short
  uint32_t m0a = padd(z0, b[0]), m0b = padd(z4, b[4]);
  m0a = paddmax(z1, b[1], m0a); m0b = paddmax(z5, b[5], m0b);
  m0a = paddmax(z2, b[2], m0a); m0b = paddmax(z6, b[6], m0b);
  m0a = paddmax(z3, b[3], m0a); m0b = paddmax(z7, b[7], m0b);
  uint32_t m1a = padd(o0, b[0]), m1b = padd(o4, b[4]);
  m1a = paddmax(o1, b[1], m1a); m1b = paddmax(o5, b[5], m1b);
  m1a = paddmax(o2, b[2], m1a); m1b = paddmax(o6, b[6], m1b);
  m1a = paddmax(o3, b[3], m1a); m1b = paddmax(o7, b[7], m1b);
  a[0] = pmax(z0, o0); a[1] = pmax(z1, o1); a[2] = pmax(z2, o2); a[3] = pmax(z3, o3);
  a[4] = pmax(z4, o4); a[5] = pmax(z5, o5); a[6] = pmax(z6, o6); a[7] = pmax(z7, o7);
  return psub(pmax(m1a, m1b), pmax(m0a, m0b));
}

/*
 * grid = (ceil(nwin_max / (4*WPJ)), n_groups), block = 128 (4 warps). Warp j of block bx handles windows
 * [(4*bx + j)*WPJ, +WPJ) of its group. Write-back (turbodecoder_iter.h:104-128 with the vec_sub / vec_lut glue folded in):
 *   MODE 0: app2[rev[i]]  = L            MODE 1: app2[rev[i]] = L - app1p[i]          MODE 2: app1p[fwd[i]] = L - app2[i]
 */
template <int MODE>
__global__ void __launch_bounds__(128) job_kernel(const Group* __restrict__ groups, const KTable* __restrict__ ktabs, uint8_t* __restrict__ ws,
                                                  const uint8_t* __restrict__ group_active, const uint8_t* __restrict__ done,
                                                  uint32_t* __restrict__ crc_acc)
{
  if (!group_active[blockIdx.y]) return;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int       wid  = threadIdx.x >> 5;
  const int       lane = threadIdx.x & 31;
  JobWarpSmem*    sm   = reinterpret_cast<JobWarpSmem*>(smem_raw) + wid;
  const Group&    g    = groups[blockIdx.y];
  const uint32_t  K    = g.K;
  const int       nwin = (int)((K + WC - 1) / WC);
  const int       w0   = (int)(blockIdx.x * 4 + wid) * WPJ;
  if (w0 >= nwin) return;
  const int       w1   = min(w0 + WPJ, nwin);
  const GroupPtrs gp   = group_ptrs(ws, g);
  const KTable&   kt   = ktabs[g.kidx];

  const uint32_t* in0  = (MODE == 2) ? gp.app2 : gp.syst;
  const uint32_t* in1  = (MODE == 2) ? gp.par1 : (MODE == 1 ? gp.app1p : gp.par0);
  const uint32_t* in2  = (MODE == 1) ? gp.par0 : nullptr;
  const uint2*    tab  = (MODE == 2) ? kt.dec2[g.crc_kind] : kt.dec1[g.crc_kind];
  uint32_t*       dst  = (MODE == 2) ? gp.app1p : gp.app2;
  uint32_t*       bits = (MODE == 2) ? gp.bits2 : gp.bits1;

  if (lane == 0) {
    mbar_init(&sm->bar[0], 1);
    mbar_init(&sm->bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int w) {
    // rows [w*WC, w*WC + WC) of each stream (streams are padded to a multiple of W >= WC rows) + the step table
    if (lane == 0) {
      JobStage& st  = sm->st[w & 1];
      uint64_t* bar = &sm->bar[w & 1];
      mbar_expect_tx(bar, (MODE == 1 ? 3u : 2u) * WC * LANES * 4 + WC * 8);
      bulk_g2s(&st.s[0][0][0], in0 + (size_t)w * WC * LANES, WC * LANES * 4, bar);
      bulk_g2s(&st.s[1][0][0], in1 + (size_t)w * WC * LANES, WC * LANES * 4, bar);
      if (MODE == 1) bulk_g2s(&st.s[2][0][0], in2 + (size_t)w * WC * LANES, WC * LANES * 4, bar);
      bulk_g2s(&st.tab[0], tab + (size_t)w * WC, WC * 8, bar);
    }
  };
  issue(w0);

  const int cb_lo = g.cb[lane], cb_hi = g.cb[32 + lane];
  uint32_t  keep  = 0;
  if (cb_lo < 0 || done[cb_lo]) keep |= 0x0000ffffu;
  if (cb_hi < 0 || done[cb_hi]) keep |= 0xffff0000u;
  uint32_t crc_lo = 0, crc_hi = 0;
  uint32_t ph0 = 0, ph1 = 0;

  for (int w = w0; w < w1; w++) {
    const int lo = w * WC;
    const int hi = (int)min((uint32_t)lo + WC, K);
    if (w + 1 < w1) {
      __syncwarp();  // every lane has finished reading the stage the next copy overwrites (window w-1)
      issue(w + 1);
    }
    uint32_t a[8], bt[8];
    {
      const uint32_t* ca = gp.ckA + (size_t)w * 8 * LANES + lane;
      const uint32_t* cb = gp.ckB + (size_t)(w + 1) * 8 * LANES + lane;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        a[i]  = ca[i * LANES];
        bt[i] = cb[i * LANES];
      }
    }
    if (w & 1) { mbar_wait(&sm->bar[1], ph1 & 1u); ph1++; } else { mbar_wait(&sm->bar[0], ph0 & 1u); ph0++; }
    const JobStage& st = sm->st[w & 1];

    // coarse pass (only for full windows): un-normalised beta[lo+8]
    uint32_t  bmid[8];
    const int nsub = (hi - lo) / 8;
    if (nsub == 2) {
#pragma unroll
      for (int i = 0; i < 8; i++) bmid[i] = bt[i];
      if ((uint32_t)hi < K) normalise(bmid);
#pragma unroll
      for (int r = 15; r >= 8; r--) {
        LOAD_XY(st, r);
        beta_step(bmid, x, y);
        if (r == 12) normalise(bmid);  // k = lo + 12
      }
    }
    uint32_t bitacc = 0;
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
      uint32_t B[8][8], cur[8];
      const bool last = (s == nsub - 1);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        cur[i]  = last ? bt[i] : bmid[i];
        B[7][i] = cur[i];
      }
      if ((uint32_t)(lo + 8 * s + 8) < K) normalise(cur);
#pragma unroll
      for (int j = 6; j >= 0; j--) {
        LOAD_XY(st, 8 * s + 1 + j);
        beta_step(cur, x, y);
#pragma unroll
        for (int i = 0; i < 8; i++) B[j][i] = cur[i];
        if (j == 3) normalise(cur);  // k = lo + 8s + 4
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int r = 8 * s + j;
        LOAD_XY(st, r);
        uint32_t ap = (MODE == 1) ? st.s[1][r][lane] : (MODE == 2 ? st.s[0][r][lane] : 0u);
        uint32_t L  = alpha_llr_step(a, B[j], x, y);
        if (j == 3 || j == 7) normalise(a);
        const uint2 t = st.tab[r];
        dst[(size_t)t.x * LANES + lane] = (MODE == 0) ? L : psub(L, ap);
        const uint32_t pos = positive_mask(L);
        crc_lo ^= (pos & 0x8000u) ? t.y : 0u;
        crc_hi ^= (pos & 0x80000000u) ? t.y : 0u;
        bitacc = ((bitacc >> 1) & 0x7fff7fffu) | pos;
      }
    }
    if (nsub == 1) bitacc = (bitacc >> 8) & 0x00ff00ffu;  // K = 8 mod 16: the last word holds eight decisions
    uint32_t* bw = bits + (size_t)w * LANES + lane;
    if (keep) bitacc = (bitacc & ~keep) | (*bw & keep);  // finished code blocks keep their final decisions
    *bw = bitacc;
  }
  if (cb_lo >= 0 && !(keep & 0xffffu) && crc_lo) atomicXor(&crc_acc[cb_lo], crc_lo);
  if (cb_hi >= 0 && !(keep & 0xffff0000u) && crc_hi) atomicXor(&crc_acc[cb_hi], crc_hi);
}
#undef LOAD_XY

/*
 * Per-code-block verdict after half-iteration number cnt (1-based): the loop condition of decode_tb_cb
 * (lib/src/phy/phch/sch.c:426-456). grid = n_groups, block = 64.
 */
__global__ void __launch_bounds__(64) status_kernel(const Group* __restrict__ groups, uint32_t* __restrict__ crc_acc, uint8_t* __restrict__ noi,
                                                    uint8_t* __restrict__ ok, uint8_t* __restrict__ done, uint8_t* __restrict__ group_active,
                                                    uint32_t cnt, uint32_t max_iter, uint32_t min_iter, int early_stop)
{
  const Group& g = groups[blockIdx.x];
  if (!group_active[blockIdx.x]) return;
  const int cb     = g.cb[threadIdx.x];
  int       active = 0;
  if (cb >= 0) {
    if (!done[cb]) {
      const uint32_t okv = (g.crc_kind != 0 && crc_acc[cb] == 0u) ? 1u : 0u;
      noi[cb] = (uint8_t)cnt;
      ok[cb]  = (uint8_t)okv;
      if ((early_stop && okv && cnt >= min_iter) || cnt >= max_iter) done[cb] = 1;
      else active = 1;
    }
    crc_acc[cb] = 0u;
  }
  active = __syncthreads_or(active);
  if (threadIdx.x == 0) group_active[blockIdx.x] = (uint8_t)(active ? 1 : 0);
}

/*
 * De-multiplex natural-order LLRs (tdec_gen_extract_input, turbodecoder_gen.c:238-258) of up to 64 code blocks into the
 * group's packed [row][lane] streams. grid = (ceil(R / 32), n_groups), block = 256.
 * llr_off[cb] = element offset of the code block's 3K+12 int16 in llr.
 */
__global__ void __launch_bounds__(256)
extract_kernel(const Group* __restrict__ groups, uint8_t* __restrict__ ws, const int16_t* __restrict__ llr, const uint64_t* __restrict__ llr_off)
{
  __shared__ int16_t tile[64][3 * 32 + 2];
  const Group&    g  = groups[blockIdx.y];
  const GroupPtrs gp = group_ptrs(ws, g);
  const uint32_t  K  = g.K;
  const uint32_t  k0 = blockIdx.x * 32;  // first row of this tile
  if (k0 >= g.R) return;
  const int tid = threadIdx.x;
  // body rows k0..k0+31 (< K): 96 consecutive int16 per code block
  const uint32_t nbody = (k0 < K) ? min(32u, K - k0) : 0u;
  for (int idx = tid; idx < 64 * 96; idx += 256) {
    int     cbl = idx / 96, e = idx % 96;
    int     cb  = g.cb[cbl];
    int16_t v   = 0;
    if (cb >= 0 && (uint32_t)e < 3 * nbody) v = llr[llr_off[cb] + 3ull * k0 + e];
    tile[cbl][e] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * 32; idx += 256) {
    int      r = idx / 32, l = idx % 32;
    uint32_t k = k0 + r;
    if (k >= g.R) continue;
    uint32_t s = 0, p0 = 0, p1 = 0, a2 = 0;
    if (k < K) {
      s  = (uint16_t)tile[l][3 * r] | ((uint32_t)(uint16_t)tile[32 + l][3 * r] << 16);
      p0 = (uint16_t)tile[l][3 * r + 1] | ((uint32_t)(uint16_t)tile[32 + l][3 * r + 1] << 16);
      p1 = (uint16_t)tile[l][3 * r + 2] | ((uint32_t)(uint16_t)tile[32 + l][3 * r + 2] << 16);
    } else if (k < K + 3) {
      // termination: syst/par0 from the first six tail values, app2/par1 from the last six
      uint32_t j = k - K;
      uint16_t v[2][4];
      for (int h = 0; h < 2; h++) {
        int cb = g.cb[32 * h + l];
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
      s  = v[0][0] | ((uint32_t)v[1][0] << 16);
      p0 = v[0][1] | ((uint32_t)v[1][1] << 16);
      a2 = v[0][2] | ((uint32_t)v[1][2] << 16);
      p1 = v[0][3] | ((uint32_t)v[1][3] << 16);
    }
    size_t o    = (size_t)k * LANES + l;
    gp.syst[o]  = s;
    gp.par0[o]  = p0;
    gp.par1[o]  = p1;
    gp.app1p[o] = 0;   // rows >= K must stay zero (no a-priori on the termination steps)
    gp.app2[o]  = a2;  // rows < K are overwritten by DEC1 before DEC2 reads them
  }
}

/*
 * Final hard-decision bytes (tdec_gen_decision_byte, MSB first) of every code block: bits of the last half-iteration
 * it ran; after a DEC2 half-iteration they are gathered through the QPP permutation (app1[fwd[i]] = ext2[i]).
 * grid = (ceil(K/32/128), n_groups*64), block = 128 : one thread per 32 output bits.
 */
__global__ void __launch_bounds__(128)
emit_kernel(const Group* __restrict__ groups, const KTable* __restrict__ ktabs, uint8_t* __restrict__ ws, const uint8_t* __restrict__ noi,
            uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off)
{
  const Group& g   = groups[blockIdx.y >> 6];
  const int    cbl = blockIdx.y & 63;
  const int    cb  = g.cb[cbl];
  if (cb < 0) return;
  const uint32_t word = blockIdx.x * 128 + threadIdx.x;  // 32-bit word of the output
  if (word * 32 >= g.K) return;
  const GroupPtrs gp   = group_ptrs(ws, g);
  const int       lane = cbl & 31, sh = (cbl >> 5) * 16;
  const uint32_t  n    = noi[cb];
  uint32_t        v    = 0;
  if (n & 1u) {
    // natural order: two 16-bit pieces
    uint32_t w0 = (gp.bits1[(size_t)(2 * word) * LANES + lane] >> sh) & 0xffffu;
    uint32_t w1 = (gp.bits1[(size_t)(2 * word + 1) * LANES + lane] >> sh) & 0xffffu;
    v           = w0 | (w1 << 16);  // bit t of v = decision of step 32*word + t
  } else {
    const uint16_t* rev = ktabs[g.kidx].rev;
    const int nt = (int)min(32u, g.K - word * 32);
    for (int t = 0; t < nt; t++) {
      uint32_t i = rev[word * 32 + t];
      uint32_t b = (gp.bits2[(size_t)(i >> 4) * LANES + lane] >> (sh + (i & 15))) & 1u;
      v |= b << t;
    }
  }
  // MSB-first bytes; K is a multiple of 8 but not always of 32
  v = __brev(v);  // bit 31 = step 32*word
  uint8_t*       o      = out + out_off[cb] + 4ull * word;
  const uint32_t nbytes = min(4u, g.K / 8 - 4 * word);
  for (uint32_t b = 0; b < nbytes; b++) o[b] = (uint8_t)(v >> (24 - 8 * b));
}

}  // namespace srsb200
