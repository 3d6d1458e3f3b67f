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
// per code block: seven int8 arrays of pitch KP = K + 16: syst, par0, par1, app1, app2, ext1, nd (natural order; syst / par0 / par1 /
// app2 carry the three termination values at K..K+2; nd = scratch for DEC2's decisions on their way back to natural order)
__host__ __device__ inline uint32_t w8_pitch(uint32_t K) { return (K + 16 + 15) & ~15u; }
__host__ __device__ inline uint64_t w8_cb_bytes(uint32_t K) { return 7ull * w8_pitch(K); }

// Shared memory of a warp: the two input streams of the constituent decode, window-interleaved. 25 KB => eight warps per SM. Everything
// else lives in registers or goes through global memory in coalesced pieces: the beta checkpoints (one 512-byte row per 8 steps,
// written and read back by the same lanes, the next one prefetched a block ahead), the a-priori values the DEC1 glue needs again at
// the output, DEC2's decisions on their way to natural order. (The first version kept all of it in shared memory - 54 KB, four
// warps per SM, one per scheduler - and ran at a quarter of an instruction per cycle and scheduler: profiles/r02_win8_ncu.md.)
struct alignas(16) W8Smem {
  uint8_t X[W8_MAX_S * W8_PITCH];  // systematic (+ a-priori) per step; overwritten step by step with the extrinsic output
  uint8_t Y[W8_MAX_S * W8_PITCH];  // parity; overwritten step by step with the hard decisions
};
constexpr size_t W8_CK_UNIT = (size_t)W8_MAX_CK * 32;  // uint4 per unit in the global checkpoint scratch

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
__device__ __forceinline__ void w8_ck_unpack(const uint4 v, uint32_t (&o)[8])
{
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
 * grid = n_units, block = 32. crcw_a / crcw_b[m] = x^(m+24) mod g (CRC24A / CRC24B), m < 6144. ckg: W8_CK_UNIT uint4 per unit.
 */
template <int MODE>
__global__ void __launch_bounds__(32) win8_kernel(const Unit8* __restrict__ units, const W8Tables* __restrict__ tabs, uint8_t* __restrict__ ws,
                                                  const uint64_t* __restrict__ ws_off, uint4* __restrict__ ckg, uint8_t* __restrict__ done,
                                                  uint8_t* __restrict__ noi, uint8_t* __restrict__ ok, uint8_t* __restrict__ out,
                                                  const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_len,
                                                  const uint32_t* __restrict__ crcw_a, const uint32_t* __restrict__ crcw_b, uint32_t cnt, uint32_t max_iter,
                                                  uint32_t min_iter, int early_stop, const uint8_t* __restrict__ max_iter_cb)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  W8Smem&        sm   = *reinterpret_cast<W8Smem*>(smem_raw);
  // (every field of the unit goes into its own register: a local copy of the struct indexed with a runtime code-block number lands
  //  in local memory, and the first version spent a quarter of its stall samples on those loads - profiles/r02_win8_ncu.md)
  const Unit8*   up   = units + blockIdx.x;
  const int      lane = threadIdx.x;
  const uint32_t K = up->K, NW = up->NW, S = up->S, KP = w8_pitch(K), crc_kind = up->crc_kind;
  const int      cb0 = up->cb[0], cb1 = up->cb[1], cb2 = up->cb[2], cb3 = up->cb[3];
  const int      npairs = 32 / (int)NW;  // 1 or 2
  const int      ncb    = 2 * npairs;
  auto cb_of = [&](int c) { return c == 0 ? cb0 : (c == 1 ? cb1 : (c == 2 ? cb2 : cb3)); };
  // ---- anything left to do?
  const bool live0 = cb0 >= 0 && !done[cb0], live1 = cb1 >= 0 && !done[cb1];
  const bool live2 = ncb > 2 && cb2 >= 0 && !done[cb2], live3 = ncb > 2 && cb3 >= 0 && !done[cb3];
  auto live_of = [&](int c) { return c == 0 ? live0 : (c == 1 ? live1 : (c == 2 ? live2 : live3)); };
  if (!(live0 || live1 || live2 || live3)) return;
  const W8Tables tb = tabs[up->kidx];
  uint4*         ck = ckg + (size_t)blockIdx.x * W8_CK_UNIT;
  // sub_glue: srsran_vec_sub_bbb saturates, except (AVX2 build) on the last K % 32 elements of the window-interleaved array, which
  // its scalar tail subtracts with wrap-around: interleaved index = step * NW + window >= K - K % 32  <=>  (K % 32 != 0 and step == S-1)
  const bool wrap_tail = (K & 31u) != 0;

  // ---------------------------------------------------------------- stage: global natural order -> shared [step][lane] byte pairs
  // Four consecutive trellis positions per lane and 32-bit global loads (the arrays are 16-byte aligned), four trips in flight:
  // every load of a batch is issued before the first value is used, so the warp pays the DRAM latency once per 512 positions.
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (c >= ncb) break;
    const int     col0 = (c >> 1) * (int)NW;  // first lane of this code block's pair
    const int     half = c & 1;
    const int     cb   = cb_of(c);
    uint8_t*      base = (cb >= 0) ? ws + ws_off[cb] : nullptr;
    constexpr int UN   = 4;
    for (uint32_t n0 = 0; n0 < K; n0 += 128 * UN) {
      uint32_t vx[UN], vy[UN], va[UN], ve[UN];
#pragma unroll
      for (int t = 0; t < UN; t++) {
        const uint32_t n = n0 + 128 * t + 4 * lane;
        vx[t] = vy[t] = va[t] = ve[t] = 0u;
        if (n < K && base) {
          // MODE 0/1: syst + par0 (+ app1, ext1); MODE 2: app2 + par1
          vx[t] = *reinterpret_cast<const uint32_t*>(base + (MODE == 2 ? 4 * KP : 0) + n);
          vy[t] = *reinterpret_cast<const uint32_t*>(base + (MODE == 2 ? 2 * KP : KP) + n);
          if (MODE == 1) {
            va[t] = *reinterpret_cast<const uint32_t*>(base + 3 * KP + n);
            ve[t] = *reinterpret_cast<const uint32_t*>(base + 5 * KP + n);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < UN; t++) {
        const uint32_t n = n0 + 128 * t + 4 * lane;
        if (n >= K) continue;
        uint32_t w = n / S, k = n - w * S, apw = 0;
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
          // (K is a multiple of 8, so the four positions exist together; they may straddle a window boundary)
          const uint32_t a = k * W8_PITCH + 2 * (col0 + w) + half;
          int x = (int)(int8_t)(vx[t] >> (8 * bb)), y = (int)(int8_t)(vy[t] >> (8 * bb));
          if (MODE == 1) {
            const int d  = (int)(int8_t)(va[t] >> (8 * bb)) - (int)(int8_t)(ve[t] >> (8 * bb));
            const int ap = (wrap_tail && k == S - 1) ? (int)(int8_t)d : w8_sat_i(d);  // app1 <- app1 - ext1 (turbodecoder_iter.h:106-108)
            x            = w8_sat_i(ap + x);                                         // simd_add(ap, x), turbodecoder_win.h:608-611
            apw |= ((uint32_t)ap & 0xffu) << (8 * bb);
          }
          sm.X[a] = (uint8_t)x;
          sm.Y[a] = (uint8_t)y;
          if (++k == S) { k = 0; w++; }
        }
        // the updated a-priori values go back where they came from: the output side subtracts them again (ext1 <- ext1 - app1)
        if (MODE == 1 && base) *reinterpret_cast<uint32_t*>(base + 3 * KP + n) = apw;
      }
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
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int st0 = 0, st1 = 0, st2 = 0, st3 = 0, st4 = 0, st5 = 0, st6 = 0, st7 = 0;
        const int cb = pair ? (h ? cb3 : cb2) : (h ? cb1 : cb0);
        if (cb >= 0) {
          const int8_t* base = reinterpret_cast<const int8_t*>(ws + ws_off[cb]);
          const int8_t* pin  = (MODE == 2) ? base + 4 * KP : base;       // app2 (second encoder's termination systematic) / syst
          const int8_t* ppa  = (MODE == 2) ? base + 2 * KP : base + KP;  // par1 / par0
          for (int k = (int)K + 2; k >= (int)K; k--) {
            const int xv = pin[k], yv = ppa[k], xy = w8_sadd_tail(xv, yv);
            const int n0 = max(w8_sadd_tail(st4, xy), st0), n1 = max(st4, w8_sadd_tail(st0, xy));
            const int n2 = max(w8_sadd_tail(st5, yv), w8_sadd_tail(st1, xv)), n3 = max(w8_sadd_tail(st5, xv), w8_sadd_tail(st1, yv));
            const int n4 = max(w8_sadd_tail(st6, xv), w8_sadd_tail(st2, yv)), n5 = max(w8_sadd_tail(st6, yv), w8_sadd_tail(st2, xv));
            const int n6 = max(st7, w8_sadd_tail(st3, xy)), n7 = max(w8_sadd_tail(st7, xy), st3);
            st0 = n0; st1 = n1; st2 = n2; st3 = n3; st4 = n4; st5 = n5; st6 = n6; st7 = n7;
          }
        }
        const int      sv[8] = {st0, st1, st2, st3, st4, st5, st6, st7};
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = h ? ((t[i] & 0xffffu) | ((uint32_t)sv[i] << 16)) : ((uint32_t)sv[i] & 0xffffu);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = t[i];
  }
  // ---------------------------------------------------------------- beta: main pass, checkpoint B[k] (before normalisation) at k = 8, 16, ... and B[S]
  const int ck_top = ((int)S + 7) / 8;  // slot of B[S]; B[8c] lives in slot c
  w8_ck_store(&ck[ck_top * 32 + lane], o);
  for (int k = (int)S - 1; k >= 0; k--) {
    w8_bstep(o, ldx(k), ldy(k));
    if (k && (k & 7) == 0) w8_ck_store(&ck[(k >> 3) * 32 + lane], o);
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
    // x[k] and y[k] have been consumed (the beta recompute of this block ran first): the extrinsic value and the decision
    // (tdec_win*_decision_byte: > 0) take their places
    *reinterpret_cast<uint16_t*>(&sm.X[k * W8_PITCH + 2 * lane]) = (uint16_t)w8_pack(ext);
    const uint32_t pos = __vadd2(__vmaxs2(ext, 0u), 0x7fff7fffu) & 0x80008000u;  // bit 15 / 31 set iff the value is > 0
    *reinterpret_cast<uint16_t*>(&sm.Y[k * W8_PITCH + 2 * lane]) = (uint16_t)(((pos >> 15) & 1u) | ((pos >> 23) & 0x100u));
  };
  uint4 ck_next = ck[(8 >= (int)S ? ck_top : 1) * 32 + lane];
  for (int k0 = 0; k0 < (int)S; k0 += 8) {
    const int len = min(8, (int)S - k0), top = k0 + len;
    // B[j] = beta[k0 + 1 + j] as stored (before normalisation): what the LLR of step k0 + j reads; B[len-1] is the checkpoint
    uint32_t B[8][8];
    uint32_t s[8];
    w8_ck_unpack(ck_next, s);
    if (top < (int)S) {  // the next block's checkpoint is fetched now and used after this block's ~1000 instructions
      const int ntop = min(top + 8, (int)S);
      ck_next        = ck[(ntop == (int)S ? ck_top : (ntop >> 3)) * 32 + lane];
    }
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
      // the short last block of a window whose length is not a multiple of 8 (top == S, at most 7 steps, once per window): beta of
      // every step is recomputed from the checkpoint on its own - no register window with a runtime length (an array indexed
      // by `len` would push the whole window into local memory, for the full blocks too)
      for (int j = 0; j < len; j++) {
        uint32_t t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = s[i];
        if (j < len - 1 && top < (int)S) w8_norm(t);
        for (int kk = top - 1; kk >= k0 + 1 + j; kk--) {
          w8_bstep(t, ldx(kk), ldy(kk));
          if (kk > k0 + 1 + j) w8_norm(t);  // kk >= 1; the last one stays as stored: before normalisation
        }
        astep_out(k0 + j, t);
      }
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- emit: natural order, QPP scatter, hard bits, CRC, verdict
  // Lanes walk consecutive trellis positions (coalesced byte accesses; the ballot of 32 decisions is one output word). Per 256
  // positions every global load (QPP table, a-priori bytes) is issued before the first value is used.
  const uint32_t nbytes = K / 8;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    if (c >= ncb) break;
    const int cb = cb_of(c);
    if (cb < 0 || !live_of(c)) continue;  // (done code blocks of a live unit are recomputed but nothing of them is stored)
    const int col0 = (c >> 1) * (int)NW, half = c & 1;
    int8_t*   base = reinterpret_cast<int8_t*>(ws + ws_off[cb]);
    int8_t *  pA1 = base + 3 * KP, *pA2 = base + 4 * KP, *pE1 = base + 5 * KP;
    uint8_t*  pND = reinterpret_cast<uint8_t*>(base + 6 * KP);
    const uint16_t* perm  = (MODE == 2) ? tb.fwd : tb.rev;
    const uint32_t* crcw  = (crc_kind == 1) ? crcw_a : crcw_b;
    uint8_t*        dst   = out + out_off[cb];
    const uint32_t  total = out_len ? out_len[cb] : nbytes;
    uint32_t        crc   = 0;
    constexpr int   UE    = 8;
    // 32 x UE decisions -> 32 output bytes: lane l takes byte l of the batch; its CRC contribution by linearity, eight consecutive
    // table words per byte (crcw[m] = x^(m+24) mod g; bit 7 of a byte = the earliest position = the highest m)
    auto finish_words = [&](const uint32_t (&bal)[UE], uint32_t n0) {
      uint32_t wv = 0;
#pragma unroll
      for (int t = 0; t < UE; t++)
        if ((lane >> 2) == t) wv = bal[t];
      const uint32_t b = (n0 >> 3) + lane;
      if (b < nbytes) {
        const uint32_t v = (__brev(wv) >> (24 - 8 * (lane & 3))) & 0xffu;  // position n in bit 31 - (n mod 32): MSB-first bytes
        if (b < total) dst[b] = (uint8_t)v;
        if (crc_kind && v) {
          const uint4* wq = reinterpret_cast<const uint4*>(crcw + (K - 8 - 8 * b));  // 32-byte aligned
          const uint4  lo = __ldg(wq), hi = __ldg(wq + 1);
          crc ^= (lo.x & (0u - (v & 1u))) ^ (lo.y & (0u - ((v >> 1) & 1u))) ^ (lo.z & (0u - ((v >> 2) & 1u))) ^ (lo.w & (0u - ((v >> 3) & 1u)));
          crc ^= (hi.x & (0u - ((v >> 4) & 1u))) ^ (hi.y & (0u - ((v >> 5) & 1u))) ^ (hi.z & (0u - ((v >> 6) & 1u))) ^ (hi.w & (0u - (v >> 7)));
        }
      }
    };
    uint32_t w = 0, k = (uint32_t)lane;  // position n = n0 + 32 t + lane walks (window, step) incrementally: S > 32
    for (uint32_t n0 = 0; n0 < K; n0 += 32 * UE) {
      uint32_t pv[UE], av[UE], bal[UE];
#pragma unroll
      for (int t = 0; t < UE; t++) {
        const uint32_t n = n0 + 32 * t + lane;
        pv[t] = (n < K) ? (uint32_t)__ldg(perm + n) : 0u;
        av[t] = (MODE == 1 && n < K) ? (uint32_t)(uint8_t)pA1[n] : 0u;
      }
#pragma unroll
      for (int t = 0; t < UE; t++) {
        const uint32_t n = n0 + 32 * t + lane;
        uint32_t       d = 0;
        if (n < K) {
          const uint32_t a = k * W8_PITCH + 2 * (col0 + w) + half;
          int            r = (int)(int8_t)sm.X[a];
          d                = sm.Y[a];
          if (MODE == 1) {
            // ext1 <- ext1 - app1 (turbodecoder_iter.h:116-118), same saturate / wrap rule as on the way in
            const int df = r - (int)(int8_t)av[t];
            r            = (wrap_tail && k == S - 1) ? (int)(int8_t)df : w8_sat_i(df);
          }
          if (MODE == 2) {
            pA1[pv[t]] = (int8_t)r;   // app1[fwd[i]] = ext2[i] (turbodecoder_iter.h:127)
            pND[pv[t]] = (uint8_t)d;  // the decision of step i belongs to natural position fwd[i]
          } else {
            pE1[n]     = (int8_t)r;   // ext1 (after the subtraction) stays for the next DEC1's glue
            pA2[pv[t]] = (int8_t)r;   // app2[rev[i]] = ext1[i] (turbodecoder_iter.h:120)
          }
        }
        k += 32;
        if (k >= S) { k -= S; w++; }
        bal[t] = __ballot_sync(0xffffffffu, d != 0);
      }
      if (MODE != 2) finish_words(bal, n0);
    }
    if (MODE == 2) {
      __syncwarp();  // the scattered decisions of all lanes are visible to all lanes
      for (uint32_t n0 = 0; n0 < K; n0 += 32 * UE) {
        uint32_t dv[UE], bal[UE];
#pragma unroll
        for (int t = 0; t < UE; t++) {
          const uint32_t n = n0 + 32 * t + lane;
          dv[t] = (n < K) ? (uint32_t)pND[n] : 0u;
        }
#pragma unroll
        for (int t = 0; t < UE; t++) bal[t] = __ballot_sync(0xffffffffu, dv[t] != 0);
        finish_words(bal, n0);
      }
    }
    crc = __reduce_xor_sync(0xffffffffu, crc);
    if (lane == 0) {
      const uint32_t okv = (crc_kind != 0 && crc == 0u) ? 1u : 0u;
      noi[cb] = (uint8_t)cnt;
      ok[cb]  = (uint8_t)okv;
      if ((early_stop && okv && cnt >= min_iter) || cnt >= (max_iter_cb ? (uint32_t)max_iter_cb[cb] : max_iter)) done[cb] = 1;
    }
    __syncwarp();
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
