/*
 * win8_kernels.cuh - the 8-BIT LLR mode on sm_100a: what srsran_tdec_iteration_8bit computes in AUTO mode
 * (lib/src/phy/fec/turbo/turbodecoder.c:458-484) for the block sizes that run in the reference's windowed saturating int8
 * decoders (lib/include/srsran/phy/fec/turbo/turbodecoder_win.h with llr_t = int8_t, :180-186 / :217-283), bit for bit:
 * NW = 32 (K > 2048, K % 32 == 0) or 16 (K > 800, K % 16 == 0) windows per code block, 40-step warm-up from the neighbouring
 * window, saturating add / subtract, max-normalisation after every step, extrinsic = (max1 - max0) >> 1. The algorithm is
 * spelled out step by step in DESIGN.md section 5b (the CPU checker of the test suite restates it and is pinned to the compiled
 * reference).
 *
 * Mapping: the windowed algorithm has no K-long dependency chain - a window is 51..192 steps - so the whole constituent decode
 * of a code block runs inside ONE warp: lane = window, the 40-step boundary states travel between lanes with warp shuffles
 * (turbodecoder_win.h moves them between SIMD lanes with byte shuffles), and two code blocks share a warp in the low / high
 * halves of every 32-bit register (int8 values in int16 lanes: VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 do the arithmetic,
 * saturation to [-128, 127] is a min / max against constants - sm_100a has no s8x4 min / max / saturating add). With 16
 * windows a warp carries four code blocks (two pairs of 16 lanes).
 *
 * Per half-iteration one launch, one warp per unit:
 *   stage   natural-order int8 streams of the unit's code blocks -> shared memory [step][lane] byte pairs (coalesced global
 *           reads, the a-priori glue of turbodecoder_iter.h:104-128 applied on the way in);
 *   beta    warm-up, shuffle, main pass storing the state every 8 steps (checkpoints, shared memory);
 *   alpha   warm-up, shuffle, main pass: per 8 steps beta is recomputed from its checkpoint into registers, then alpha + LLR;
 *           extrinsic bytes and hard decisions overwrite the consumed a-priori / parity slots;
 *   emit    extrinsic back to natural order in global memory + QPP scatter for the next half-iteration, hard bits by warp
 *           ballot (32 trellis steps per word), CRC24 by linearity, per-code-block verdict (sch.c:426-456), decoded bytes.
 * Units whose code blocks are all done return at once: early termination works per warp, not per 64-block group.
 * HBM traffic per code-block step and half-iteration: 3-4 bytes in, 2 out - the kernel is bound by integer issue.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

constexpr int      W8_OVERLAP  = 40;   // win_overlap_len
constexpr int      W8_MAX_S    = 192;  // longest window: 6144 / 32
constexpr int      W8_PITCH    = 66;   // bytes per shared-memory row of 32 byte pairs (+2: the transposing stores of the staging spread over all banks)
constexpr int      W8_MAX_CK   = W8_MAX_S / 8 + 2;
constexpr uint32_t W8_M128     = 0xFF80FF80u;  // two int16 of -128
constexpr uint32_t W8_P127     = 0x007F007Fu;  // two int16 of +127

struct Unit8 {
  uint32_t K, NW, S, kidx, crc_kind;  // NW windows of S steps; crc_kind 0 none, 1 CRC24A, 2 CRC24B
  int32_t  cb[4];                     // code-block ids: pair p = lanes [p*NW, (p+1)*NW), cb[2p] low half, cb[2p+1] high half; -1 = empty
  uint32_t pad_;
};
// per code block: six int8 arrays of pitch KP = K + 16: syst, par0, par1, app1, app2, ext1 (natural order; syst / par0 / par1 /
// app2 carry the three termination values at K..K+2)
__host__ __device__ inline uint32_t w8_pitch(uint32_t K) { return (K + 16 + 15) & ~15u; }
__host__ __device__ inline uint64_t w8_cb_bytes(uint32_t K) { return 6ull * w8_pitch(K); }

struct alignas(16) W8Smem {
  uint8_t  X[W8_MAX_S * W8_PITCH];     // systematic (+ a-priori) per step; scratch for the natural-order decisions afterwards
  uint8_t  Y[W8_MAX_S * W8_PITCH];     // parity; hard decisions after the alpha pass
  uint8_t  A[W8_MAX_S * W8_PITCH];     // a-priori (DEC1 with a-priori); extrinsic output after the alpha pass
  uint4    ck[W8_MAX_CK][32];          // beta checkpoints: 8 states x 2 code blocks as bytes
  uint32_t hard[4][W8_MAX_S];          // decoded bits of each of the unit's code blocks, 32 steps per word (MSB first)
};

// ---- int8 values in int16x2 lanes
// PRMT with the sign-replicating selectors (nibble 8 | n = the sign of byte n in all eight bits). Inline PTX on purpose: the
// __byte_perm intrinsic masks the selector nibbles to three bits (0x9180 becomes 0x1100 in the SASS).
__device__ __forceinline__ uint32_t w8_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t w8_unpack(uint32_t pair) { return w8_prmt(pair, 0u, 0x9180u); }  // bytes (lo, hi) -> two sign-extended int16
__device__ __forceinline__ uint32_t w8_pack(uint32_t v) { return __byte_perm(v, 0u, 0x4420u) & 0xffffu; }
__device__ __forceinline__ uint32_t w8_sat(uint32_t v) { return __vmins2(__vmaxs2(v, W8_M128), W8_P127); }
__device__ __forceinline__ uint32_t w8_adds(uint32_t a, uint32_t b) { return __vmins2(__viaddmax_s16x2(a, b, W8_M128), W8_P127); }
__device__ __forceinline__ uint32_t w8_subs(uint32_t a, uint32_t b) { return w8_sat(__vsub2(a, b)); }
// max(sat(a + b), c) for c already in range: the lower clamp is implied by c
__device__ __forceinline__ uint32_t w8_addmax_hi(uint32_t a, uint32_t b, uint32_t c) { return __vmins2(__viaddmax_s16x2(a, b, c), W8_P127); }

__device__ __forceinline__ void w8_norm(uint32_t (&o)[8])
{
  // normalize_max (turbodecoder_win.h:479-497): subtract the maximum, saturating (the difference is <= 0: only the lower clamp)
  const uint32_t m  = __vmaxs2(__vimax3_s16x2(__vimax3_s16x2(__vimax3_s16x2(o[0], o[1], o[2]), o[3], o[4]), o[5], o[6]), o[7]);
  const uint32_t nm = __vsub2(0u, m);
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = __viaddmax_s16x2(o[i], nm, W8_M128);
}
// backward step, turbodecoder_win.h:613-636
__device__ __forceinline__ void w8_bstep(uint32_t (&o)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = w8_adds(x, y);
  const uint32_t n0 = w8_addmax_hi(o[4], xy, o[0]);
  const uint32_t n1 = w8_addmax_hi(o[0], xy, o[4]);
  const uint32_t n6 = w8_addmax_hi(o[3], xy, o[7]);
  const uint32_t n7 = w8_addmax_hi(o[7], xy, o[3]);
  const uint32_t n2 = w8_sat(__viaddmax_s16x2(o[5], y, __vadd2(o[1], x)));
  const uint32_t n3 = w8_sat(__viaddmax_s16x2(o[5], x, __vadd2(o[1], y)));
  const uint32_t n4 = w8_sat(__viaddmax_s16x2(o[6], x, __vadd2(o[2], y)));
  const uint32_t n5 = w8_sat(__viaddmax_s16x2(o[6], y, __vadd2(o[2], x)));
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}
// branch sums of a forward step, turbodecoder_win.h:724-741: z = information bit 0 into state i, w = information bit 1
__device__ __forceinline__ void w8_abranches(const uint32_t (&o)[8], uint32_t x, uint32_t y, uint32_t (&z)[8], uint32_t (&w)[8])
{
  const uint32_t xy = w8_adds(x, y);
  z[0] = o[0]; z[1] = w8_adds(o[3], y); z[2] = w8_adds(o[4], y); z[3] = o[7];
  z[4] = o[1]; z[5] = w8_adds(o[2], y); z[6] = w8_adds(o[5], y); z[7] = o[6];
  w[0] = w8_adds(o[1], xy); w[1] = w8_adds(o[2], x); w[2] = w8_adds(o[5], x); w[3] = w8_adds(o[6], xy);
  w[4] = w8_adds(o[0], xy); w[5] = w8_adds(o[3], x); w[6] = w8_adds(o[4], x); w[7] = w8_adds(o[7], xy);
}
// forward step without output (warm-up): max(sat(u), sat(v)) = sat(max(u, v))
__device__ __forceinline__ void w8_astep(uint32_t (&o)[8], uint32_t x, uint32_t y)
{
  const uint32_t xy = w8_adds(x, y);
  const uint32_t n0 = w8_addmax_hi(o[1], xy, o[0]);
  const uint32_t n3 = w8_addmax_hi(o[6], xy, o[7]);
  const uint32_t n4 = w8_addmax_hi(o[0], xy, o[1]);
  const uint32_t n7 = w8_addmax_hi(o[7], xy, o[6]);
  const uint32_t n1 = w8_sat(__viaddmax_s16x2(o[3], y, __vadd2(o[2], x)));
  const uint32_t n2 = w8_sat(__viaddmax_s16x2(o[4], y, __vadd2(o[5], x)));
  const uint32_t n5 = w8_sat(__viaddmax_s16x2(o[2], y, __vadd2(o[3], x)));
  const uint32_t n6 = w8_sat(__viaddmax_s16x2(o[5], y, __vadd2(o[4], x)));
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

// scalar helpers of the staging / termination code
__device__ __forceinline__ int w8_sat_i(int v) { return min(127, max(-128, v)); }
// the termination steps' helper: saturates upwards only, wraps below -128 (turbodecoder_win.h:469-477)
__device__ __forceinline__ int w8_sadd_tail(int a, int b)
{
  const int z = a + b;
  return z > 127 ? 127 : (int)(int8_t)z;
}

__device__ __forceinline__ void w8_ck_store(uint4* dst, const uint32_t (&o)[8])
{
  *dst = make_uint4(__byte_perm(o[0], o[1], 0x6420u), __byte_perm(o[2], o[3], 0x6420u), __byte_perm(o[4], o[5], 0x6420u), __byte_perm(o[6], o[7], 0x6420u));
}
__device__ __forceinline__ void w8_ck_load(const uint4* src, uint32_t (&o)[8])
{
  const uint4 v = *src;
  o[0] = w8_prmt(v.x, 0u, 0x9180u); o[1] = w8_prmt(v.x, 0u, 0xB3A2u);
  o[2] = w8_prmt(v.y, 0u, 0x9180u); o[3] = w8_prmt(v.y, 0u, 0xB3A2u);
  o[4] = w8_prmt(v.z, 0u, 0x9180u); o[5] = w8_prmt(v.z, 0u, 0xB3A2u);
  o[6] = w8_prmt(v.w, 0u, 0x9180u); o[7] = w8_prmt(v.w, 0u, 0xB3A2u);
}

struct W8Tables {
  const uint16_t* fwd;   // QPP pi(i)
  const uint16_t* rev;   // inverse
};

/*
 * MODE 0: DEC1, first half-iteration     x = syst                         y = par0
 * MODE 1: DEC1 with a-priori             a = app1 - ext1 (glue), x = a + syst   y = par0
 * MODE 2: DEC2                           x = app2                         y = par1
 * grid = n_units, block = 32. crcw[kind][m] = x^(m+24) mod g (kind 1: CRC24A, 2: CRC24B), m < 6144.
 */
template <int MODE>
__global__ void __launch_bounds__(32) win8_kernel(const Unit8* __restrict__ units, const W8Tables* __restrict__ tabs, uint8_t* __restrict__ ws,
                                                  const uint64_t* __restrict__ ws_off, uint8_t* __restrict__ done, uint8_t* __restrict__ noi,
                                                  uint8_t* __restrict__ ok, uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off,
                                                  const uint32_t* __restrict__ out_len, const uint32_t* __restrict__ crcw_a,
                                                  const uint32_t* __restrict__ crcw_b, uint32_t cnt, uint32_t max_iter, uint32_t min_iter, int early_stop,
                                                  const uint8_t* __restrict__ max_iter_cb)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  W8Smem&        sm   = *reinterpret_cast<W8Smem*>(smem_raw);
  const Unit8    u    = units[blockIdx.x];
  const int      lane = threadIdx.x;
  const uint32_t K = u.K, NW = u.NW, S = u.S, KP = w8_pitch(K);
  const int      npairs = 32 / (int)NW;  // 1 or 2
  const int      ncb    = 2 * npairs;
  // ---- anything left to do?
  bool live[4];
  bool any = false;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    live[c] = c < ncb && u.cb[c] >= 0 && !done[u.cb[c]];
    any |= live[c];
  }
  if (!any) return;
  const W8Tables tb = tabs[u.kidx];
  // sub_glue: srsran_vec_sub_bbb saturates, except (AVX2 build) on the last K % 32 elements of the window-interleaved array, which
  // its scalar tail subtracts with wrap-around: interleaved index = step * NW + window >= K - K % 32  <=>  (K % 32 != 0 and step == S-1)
  const bool wrap_tail = (K & 31u) != 0;

  // ---------------------------------------------------------------- stage: global natural order -> shared [step][lane] byte pairs
  for (int c = 0; c < ncb; c++) {
    const int      col0 = (c >> 1) * (int)NW;  // first lane of this code block's pair
    const int      half = c & 1;
    const int      cb   = u.cb[c];
    if (cb < 0) {
      for (uint32_t n = lane; n < K; n += 32) {
        const uint32_t w = n / S, k = n - w * S, a = k * W8_PITCH + 2 * (col0 + w) + half;
        sm.X[a] = 0; sm.Y[a] = 0; sm.A[a] = 0;
      }
      continue;
    }
    const int8_t* base = reinterpret_cast<const int8_t*>(ws + ws_off[cb]);
    const int8_t *pS = base, *pP0 = base + KP, *pP1 = base + 2 * KP, *pA1 = base + 3 * KP, *pA2 = base + 4 * KP, *pE1 = base + 5 * KP;
    for (uint32_t n = lane; n < K; n += 32) {
      const uint32_t w = n / S, k = n - w * S, a = k * W8_PITCH + 2 * (col0 + w) + half;
      int x, y, ap = 0;
      if (MODE == 0) {
        x = pS[n]; y = pP0[n];
      } else if (MODE == 1) {
        const int d = (int)pA1[n] - (int)pE1[n];
        ap = (wrap_tail && k == S - 1) ? (int)(int8_t)d : w8_sat_i(d);  // app1 <- app1 - ext1 (turbodecoder_iter.h:106-108)
        x  = w8_sat_i(ap + (int)pS[n]);                                // simd_add(ap, x), turbodecoder_win.h:608-611
        y  = pP0[n];
      } else {
        x = pA2[n]; y = pP1[n];
      }
      sm.X[a] = (uint8_t)x; sm.Y[a] = (uint8_t)y;
      if (MODE == 1) sm.A[a] = (uint8_t)ap;
    }
  }
  __syncwarp();
  const int w_lane = lane % (int)NW;  // window of this lane within its pair
  const int pair   = lane / (int)NW;
  auto ldx = [&](int k) { return w8_unpack(*reinterpret_cast<const uint16_t*>(&sm.X[k * W8_PITCH + 2 * lane])); };
  auto ldy = [&](int k) { return w8_unpack(*reinterpret_cast<const uint16_t*>(&sm.Y[k * W8_PITCH + 2 * lane])); };

  uint32_t o[8];
  // ---------------------------------------------------------------- beta: warm-up over the window's own first 40 steps
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = 0u;  // simd_set1(-INF), INF = 0
  for (int k = W8_OVERLAP - 1; k >= 0; k--) {
    w8_bstep(o, ldx(k), ldy(k));
    if (k) w8_norm(o);
  }
  // hand the state to the window on the left; the last window starts from the termination steps (beta_trellis, :499-549)
  {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = __shfl_down_sync(0xffffffffu, o[i], 1, (int)NW);
    if (w_lane == (int)NW - 1) {
      int st[2][8];
      for (int h = 0; h < 2; h++) {
        const int cb = u.cb[2 * pair + h];
#pragma unroll
        for (int i = 0; i < 8; i++) st[h][i] = 0;
        if (cb < 0) continue;
        const int8_t* base = reinterpret_cast<const int8_t*>(ws + ws_off[cb]);
        const int8_t* pin  = (MODE == 2) ? base + 4 * KP : base;           // app2 (second encoder's termination systematic) / syst
        const int8_t* ppa  = (MODE == 2) ? base + 2 * KP : base + KP;      // par1 / par0
        for (int k = (int)K + 2; k >= (int)K; k--) {
          const int xv = pin[k], yv = ppa[k], xy = w8_sadd_tail(xv, yv);
          const int a[8] = {w8_sadd_tail(st[h][4], xy), st[h][4], w8_sadd_tail(st[h][5], yv), w8_sadd_tail(st[h][5], xv),
                            w8_sadd_tail(st[h][6], xv), w8_sadd_tail(st[h][6], yv), st[h][7], w8_sadd_tail(st[h][7], xy)};
          const int b[8] = {st[h][0], w8_sadd_tail(st[h][0], xy), w8_sadd_tail(st[h][1], xv), w8_sadd_tail(st[h][1], yv),
                            w8_sadd_tail(st[h][2], yv), w8_sadd_tail(st[h][2], xv), w8_sadd_tail(st[h][3], xy), st[h][3]};
#pragma unroll
          for (int i = 0; i < 8; i++) st[h][i] = max(a[i], b[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i++) t[i] = ((uint32_t)st[0][i] & 0xffffu) | ((uint32_t)st[1][i] << 16);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = t[i];
  }
  // ---------------------------------------------------------------- beta: main pass, checkpoint B[k] (before normalisation) at k = 8, 16, ... and B[S]
  const int ck_top = ((int)S + 7) / 8;  // slot of B[S]; B[8c] lives in slot c
  w8_ck_store(&sm.ck[ck_top][lane], o);
  for (int k = (int)S - 1; k >= 0; k--) {
    w8_bstep(o, ldx(k), ldy(k));
    if (k && (k & 7) == 0) w8_ck_store(&sm.ck[k >> 3][lane], o);
    if (k) w8_norm(o);
  }
  // ---------------------------------------------------------------- alpha: warm-up over the window's own last 40 steps
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = 0u;
  for (int j = 0; j < W8_OVERLAP; j++) {
    const int k = (int)S - W8_OVERLAP + j;
    w8_astep(o, ldx(k), ldy(k));
    if (j) w8_norm(o);
  }
  {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = __shfl_up_sync(0xffffffffu, o[i], 1, (int)NW);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = (w_lane == 0) ? 0u : t[i];  // first window: the known state {0, -INF x 7}, INF = 0
  }
  // ---------------------------------------------------------------- alpha: main pass, 8 steps at a time with beta recomputed into registers
  __syncwarp();
  // one step: branch sums, LLR against beta[k+1] (b), extrinsic / decision into the consumed slots, state update
  auto astep_out = [&](int k, const uint32_t (&b)[8]) {
    const uint32_t x = ldx(k), y = ldy(k);
    uint32_t       z[8], w[8];
    w8_abranches(o, x, y, z, w);
    // max_i sat(b_i + z_i) = sat(max_i (b_i + z_i)): saturation is monotone
    const uint32_t p0 = __vadd2(b[0], z[0]), p1 = __vadd2(b[1], z[1]), p2 = __vadd2(b[2], z[2]), p3 = __vadd2(b[3], z[3]);
    const uint32_t p4 = __vadd2(b[4], z[4]), p5 = __vadd2(b[5], z[5]), p6 = __vadd2(b[6], z[6]), p7 = __vadd2(b[7], z[7]);
    const uint32_t q0 = __vadd2(b[0], w[0]), q1 = __vadd2(b[1], w[1]), q2 = __vadd2(b[2], w[2]), q3 = __vadd2(b[3], w[3]);
    const uint32_t q4 = __vadd2(b[4], w[4]), q5 = __vadd2(b[5], w[5]), q6 = __vadd2(b[6], w[6]), q7 = __vadd2(b[7], w[7]);
    const uint32_t m0 = w8_sat(__vmaxs2(__vimax3_s16x2(__vimax3_s16x2(p0, p1, p2), p6, p7), __vimax3_s16x2(p3, p4, p5)));
    const uint32_t m1 = w8_sat(__vmaxs2(__vimax3_s16x2(__vimax3_s16x2(q0, q1, q2), q6, q7), __vimax3_s16x2(q3, q4, q5)));
    const uint32_t l  = w8_subs(m1, m0);
    // out = l >> 1, arithmetic, per element (simd_rb_shift, divide_output = 1)
    const uint32_t ext = ((l >> 1) & 0x7fff7fffu) | (l & 0x80008000u);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = __vmaxs2(z[i], w[i]);
    if (k) w8_norm(o);
    // what the next half-iteration reads, and the decision (tdec_win*_decision_byte: > 0) of this one
    uint32_t r = ext;
    if (MODE == 1) {
      const uint32_t ap = w8_unpack(*reinterpret_cast<const uint16_t*>(&sm.A[k * W8_PITCH + 2 * lane]));
      // ext1 <- ext1 - app1 (turbodecoder_iter.h:116-118), same saturate / wrap rule as on the way in
      r = (wrap_tail && k == (int)S - 1) ? __vsub2(ext, ap) : w8_subs(ext, ap);
    }
    *reinterpret_cast<uint16_t*>(&sm.A[k * W8_PITCH + 2 * lane]) = (uint16_t)w8_pack(r);
    const uint32_t pos = __vadd2(__vmaxs2(ext, 0u), 0x7fff7fffu) & 0x80008000u;  // bit 15 / 31 set iff the value is > 0
    *reinterpret_cast<uint16_t*>(&sm.Y[k * W8_PITCH + 2 * lane]) = (uint16_t)(((pos >> 15) & 1u) | ((pos >> 23) & 0x100u));
  };
  for (int k0 = 0; k0 < (int)S; k0 += 8) {
    const int len = min(8, (int)S - k0), top = k0 + len;
    // B[j] = beta[k0 + 1 + j] as stored (before normalisation): what the LLR of step k0 + j reads; B[len-1] is the checkpoint
    uint32_t B[8][8];
    uint32_t s[8];
    w8_ck_load(&sm.ck[top == (int)S ? ck_top : (top >> 3)][lane], s);
    if (len == 8) {
#pragma unroll
      for (int i = 0; i < 8; i++) B[7][i] = s[i];
      if (top < (int)S) w8_norm(s);  // the recursion went on from the normalised state; the start state B[S] is used as it is
#pragma unroll
      for (int j = 6; j >= 0; j--) {
        const int kk = k0 + 1 + j;  // >= 1: always normalised afterwards
        w8_bstep(s, ldx(kk), ldy(kk));
#pragma unroll
        for (int i = 0; i < 8; i++) B[j][i] = s[i];
        w8_norm(s);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) astep_out(k0 + j, B[j]);
    } else {
      // the short last block of a window whose length is not a multiple of 8 (top == S): same recursion, predicated
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j == len - 1) {
#pragma unroll
          for (int i = 0; i < 8; i++) B[j][i] = s[i];
        }
      if (top < (int)S) w8_norm(s);
#pragma unroll
      for (int j = 6; j >= 0; j--) {
        if (j <= len - 2) {
          const int kk = k0 + 1 + j;
          w8_bstep(s, ldx(kk), ldy(kk));
#pragma unroll
          for (int i = 0; i < 8; i++) B[j][i] = s[i];
          w8_norm(s);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < len) astep_out(k0 + j, B[j]);
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- emit: natural order, QPP scatter, hard bits, CRC, verdict
  const uint32_t nwords = (K + 31) / 32;
  for (int c = 0; c < ncb; c++) {
    const int cb = u.cb[c];
    if (cb < 0) continue;  // (done code blocks of a live unit are recomputed but nothing of them is stored)
    const int col0 = (c >> 1) * (int)NW, half = c & 1;
    int8_t*   base = reinterpret_cast<int8_t*>(ws + ws_off[cb]);
    int8_t *  pA1 = base + 3 * KP, *pA2 = base + 4 * KP, *pE1 = base + 5 * KP;
    uint8_t*  natd = sm.X;  // decisions in natural order (DEC2 produces them in interleaved order)
    if (!live[c]) continue;
    if (MODE == 2) {
      for (uint32_t n = lane; n < K; n += 32) {
        const uint32_t w = n / S, k = n - w * S, a = k * W8_PITCH + 2 * (col0 + w) + half;
        const uint32_t f = tb.fwd[n];
        pA1[f]  = (int8_t)sm.A[a];  // app1[fwd[i]] = ext2[i] (turbodecoder_iter.h:127)
        natd[f] = sm.Y[a];
      }
      __syncwarp();
    }
    uint32_t crc = 0;
    const uint32_t* crcw = (u.crc_kind == 1) ? crcw_a : crcw_b;
    for (uint32_t n0 = 0; n0 < K; n0 += 32) {
      const uint32_t n = n0 + lane;
      uint32_t       d = 0;
      if (n < K) {
        if (MODE == 2) {
          d = natd[n];
        } else {
          const uint32_t w = n / S, k = n - w * S, a = k * W8_PITCH + 2 * (col0 + w) + half;
          const int8_t   r = (int8_t)sm.A[a];
          pE1[n]           = r;            // ext1 (after the subtraction) stays for the next DEC1's glue
          pA2[tb.rev[n]]   = r;            // app2[rev[i]] = ext1[i] (turbodecoder_iter.h:120)
          d                = sm.Y[a];
        }
        if (d && u.crc_kind) crc ^= __ldg(&crcw[K - 1 - n]);
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, d != 0);
      if (lane == 0) sm.hard[c][n0 >> 5] = __brev(bal);  // step n0 in bit 31: MSB-first bytes once stored big-endian
    }
    crc = __reduce_xor_sync(0xffffffffu, crc);
    __syncwarp();
    // decoded bytes (every half-iteration: the latest decision is what a finished block keeps)
    {
      uint8_t*       dst   = out + out_off[cb];
      const uint32_t total = out_len ? out_len[cb] : K / 8;
      for (uint32_t b = lane; b < total; b += 32) dst[b] = (uint8_t)(sm.hard[c][b >> 2] >> (24 - 8 * (b & 3u)));
    }
    if (lane == 0) {
      const uint32_t okv = (u.crc_kind != 0 && crc == 0u) ? 1u : 0u;
      noi[cb] = (uint8_t)cnt;
      ok[cb]  = (uint8_t)okv;
      if ((early_stop && okv && cnt >= min_iter) || cnt >= (max_iter_cb ? (uint32_t)max_iter_cb[cb] : max_iter)) done[cb] = 1;
    }
    __syncwarp();
    (void)nwords;
  }
}

/*
 * De-multiplex natural-order int8 LLRs (s p p' triples + 12 termination values: tdec_win*_extract_input, turbodecoder_win.h:
 * 883-921, in natural order) into the per-code-block arrays and re-arm the decode state. grid = (ceil(K_max / 256), n_cb).
 */
__global__ void __launch_bounds__(256) extract8_kernel(const int8_t* __restrict__ llr, const uint64_t* __restrict__ llr_off, const uint32_t* __restrict__ cbK,
                                                       uint8_t* __restrict__ ws, const uint64_t* __restrict__ ws_off, uint8_t* __restrict__ done)
{
  const uint32_t cb = blockIdx.y, K = cbK[cb], KP = w8_pitch(K);
  const int8_t*  in = llr + llr_off[cb];
  int8_t*        b  = reinterpret_cast<int8_t*>(ws + ws_off[cb]);
  const uint32_t n  = blockIdx.x * 256 + threadIdx.x;
  if (n < K) {
    b[n]          = in[3 * n];
    b[KP + n]     = in[3 * n + 1];
    b[2 * KP + n] = in[3 * n + 2];
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < 3) {
      const uint32_t j = threadIdx.x;
      b[K + j]          = in[3 * K + 2 * j];          // syst termination
      b[KP + K + j]     = in[3 * K + 2 * j + 1];      // par0
      b[4 * KP + K + j] = in[3 * K + 6 + 2 * j];      // app2: second encoder's termination systematic
      b[2 * KP + K + j] = in[3 * K + 6 + 2 * j + 1];  // par1
    }
    if (threadIdx.x == 0) done[cb] = 0;
  }
}

/*
 * srsran_rm_turbo_rx_lut_8bit (rm_turbo.c:447-483): output[T[i mod L]] += input[i] in WRAPPING int8, natural layout, gather form
 * like rm_rx_kernel: soft-buffer position p collects e[Tinv[p] + m L]. One block per code block.
 */
struct RmJob8 {
  const int8_t*   e;
  int8_t*         buf;
  const uint16_t* table;  // Tinv
  uint32_t        E, L;
};
__global__ void __launch_bounds__(256) rm_rx8_kernel(const RmJob8* __restrict__ jobs)
{
  const RmJob8 j = jobs[blockIdx.x];
  for (uint32_t p = threadIdx.x; p < j.L; p += 256) {
    const uint32_t t = j.table[p];
    int            acc = 0;
    for (uint32_t i = t; i < j.E; i += j.L) acc += j.e[i];
    if (t < j.E) j.buf[p] = (int8_t)((int)j.buf[p] + acc);
  }
}

}  // namespace srsb200
