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
 * Data layout = compute layout. The reference keeps its 8-bit arrays window-interleaved (element of window w, step k at index
 * k * NW + w: tdec_win*_extract_input, turbodecoder_win.h:883-921) and so does this engine, with the two code blocks of a pair
 * packed byte by byte: every stream of a warp's UNIT is a [step][32 lanes] array of byte pairs, 64 bytes per step, in HBM
 * exactly as in shared memory. Staging a stream is a straight 128-bit copy, every per-step access of the decode loop is one
 * coalesced 64-byte row, the QPP interleaver is a table of destination offsets in the same layout (shared by both blocks of a
 * pair: one 2-byte scatter store moves both), and the CRC is a table of per-position remainders in that order. The first
 * version kept natural-order arrays and transposed in and out of shared memory every half-iteration: the transposing emit pass
 * cost as many instructions as the decode itself (profiles/r02_win8_ncu.md).
 *
 * Per half-iteration one launch, one warp per unit:
 *   stage   x and y streams -> shared memory (DEC1 with a-priori: x = sat(sat(app1 - ext1) + syst), turbodecoder_iter.h:104-128);
 *   beta    warm-up, shuffle, main pass with the state checkpointed every 8 steps (global scratch, coalesced, read back by the
 *           same lanes one block ahead of use);
 *   alpha   warm-up, shuffle, main pass: per 8 steps beta is recomputed from its checkpoint into registers, then alpha + LLR,
 *           and every step stores its extrinsic value in place and through the QPP table, its hard decision, and folds its
 *           CRC contribution; the per-code-block verdict (sch.c:426-456) closes the launch.
 * Units whose code blocks are all done return at once: early termination works per warp, not per 64-block group.
 * HBM traffic per code-block step and half-iteration: 2-4 bytes in, 3 out - the kernel is bound by integer issue.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

constexpr int      W8_OVERLAP = 40;   // win_overlap_len
constexpr int      W8_MAX_S   = 192;  // longest window: 6144 / 32
constexpr int      W8_ROW     = 64;   // bytes per step row: 32 lanes x 2 code blocks
constexpr int      W8_MAX_CK  = W8_MAX_S / 8 + 2;
constexpr uint32_t W8_M128    = 0xFF80FF80u;  // two int16 of -128
constexpr uint32_t W8_P127    = 0x007F007Fu;  // two int16 of +127

struct Unit8 {
  uint32_t K, NW, S, kidx, crc_kind;  // NW windows of S steps; crc_kind 0 none, 1 CRC24A, 2 CRC24B
  int32_t  cb[4];                     // code-block ids: pair p = lanes [p*NW, (p+1)*NW), cb[2p] low byte, cb[2p+1] high byte; -1 = empty
  uint32_t pad_;
  uint64_t ws_off;                    // byte offset of the unit's arrays
};
// unit workspace: seven [S][64] arrays - syst, par0, par1, app1, app2, ext1, nd (hard decisions at their natural position) - and a
// 64-byte block with the 3 termination values of {syst, par0, app2, par1} per code block: tail[cb slot][stream][4]
enum { W8_SYST = 0, W8_PAR0, W8_PAR1, W8_APP1, W8_APP2, W8_EXT1, W8_ND, W8_NARR };
__host__ __device__ inline uint64_t w8_arr_bytes(uint32_t S) { return (uint64_t)S * W8_ROW; }
__host__ __device__ inline uint64_t w8_unit_bytes(uint32_t S) { return W8_NARR * w8_arr_bytes(S) + 64; }

// Shared memory of a warp: the two input streams of the constituent decode. 24 KB => nine warps per SM.
struct alignas(16) W8Smem {
  uint8_t X[W8_MAX_S * W8_ROW];  // systematic (+ a-priori) per step
  uint8_t Y[W8_MAX_S * W8_ROW];  // parity
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
__device__ __forceinline__ uint32_t w8_unpack_hi(uint32_t w) { return w8_prmt(w, 0u, 0xB3A2u); }     // the same for bytes 2, 3 of a word
__device__ __forceinline__ uint32_t w8_pack(uint32_t v) { return __byte_perm(v, 0u, 0x4420u); }      // two int16 -> bytes (lo, hi), upper half zero
__device__ __forceinline__ uint32_t w8_sat(uint32_t v) { return __vmins2(__vmaxs2(v, W8_M128), W8_P127); }
__device__ __forceinline__ uint32_t w8_adds(uint32_t a, uint32_t b) { return __vmins2(__viaddmax_s16x2(a, b, W8_M128), W8_P127); }
__device__ __forceinline__ uint32_t w8_subs(uint32_t a, uint32_t b) { return w8_sat(__vsub2(a, b)); }
// int8 wrap-around of a 16-bit lane value (the scalar tail of srsran_vec_sub_bbb): sign-extend the low byte of each half
__device__ __forceinline__ uint32_t w8_wrap8(uint32_t v) { return w8_prmt(v, 0u, 0xA280u); }
// max(sat(a + b), c) for c already in range: the lower clamp is implied by c
__device__ __forceinline__ uint32_t w8_addmax_hi(uint32_t a, uint32_t b, uint32_t c) { return __vmins2(__viaddmax_s16x2(a, b, c), W8_P127); }

/*
 * State metrics live NEGATED AND BIASED: m" = 127 - m, a value in [0, 255] held in an int16 lane. The map is a monotone
 * bijection, so every result is the same bit for bit, and
 *   max becomes min;   the upper saturation bound +127 becomes 0 - the .RELU of VIMNMX / VIMNMX3 / VIADDMNMX -
 *   and the lower bound -128 becomes a min against 255, which is the third operand of those instructions or implied by a
 *   min against a value that is in range already.
 * A saturating add of a branch metric g is ONE instruction, max(min(m" + (-g), 255), 0), instead of two; a state update
 * max(sat(a + g), c) is one, max(min(a" + (-g), c"), 0). The branch metrics arrive negated: shared memory holds ~x and ~y.
 * First version (values as they are): 146 min / max-pipe instructions per window step, this one 101.
 */
constexpr uint32_t W8_C255 = 0x00FF00FFu, W8_C128 = 0x00800080u, W8_M127 = 0xFF81FF81u, W8_ONE = 0x00010001u;
// max(sat(a + g), c)  ->  min-form with ng = -g, c" in range
__device__ __forceinline__ uint32_t w8n_addmin(uint32_t a, uint32_t ng, uint32_t c) { return __viaddmin_s16x2_relu(a, ng, c); }
// sat(a + g)
__device__ __forceinline__ uint32_t w8n_adds(uint32_t a, uint32_t ng) { return __viaddmin_s16x2_relu(a, ng, W8_C255); }
// -sat(x + y) from nx = -x, ny = -y
__device__ __forceinline__ uint32_t w8n_nxy(uint32_t nx, uint32_t ny) { return __vmins2(__viaddmax_s16x2(nx, ny, W8_M127), W8_C128); }

__device__ __forceinline__ void w8_norm(uint32_t (&o)[8])
{
  // normalize_max (turbodecoder_win.h:479-497): subtract the maximum, saturating. Here: add 127 - min", clamp at 255
  const uint32_t m = __vmins2(__vimin3_s16x2(__vimin3_s16x2(__vimin3_s16x2(o[0], o[1], o[2]), o[3], o[4]), o[5], o[6]), o[7]);
  const uint32_t c = __vsub2(W8_P127, m);
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = __viaddmin_s16x2(o[i], c, W8_C255);
}
// backward step, turbodecoder_win.h:613-636 (nx = -x, ny = -y)
__device__ __forceinline__ void w8_bstep(uint32_t (&o)[8], uint32_t nx, uint32_t ny)
{
  const uint32_t nxy = w8n_nxy(nx, ny);
  const uint32_t n0 = w8n_addmin(o[4], nxy, o[0]);
  const uint32_t n1 = w8n_addmin(o[0], nxy, o[4]);
  const uint32_t n6 = w8n_addmin(o[3], nxy, o[7]);
  const uint32_t n7 = w8n_addmin(o[7], nxy, o[3]);
  // sat(max(a + g, b + h)) = max(min(a" - g, b" - h, 255), 0): saturation is monotone
  const uint32_t n2 = __vimin3_s16x2_relu(__vadd2(o[5], ny), __vadd2(o[1], nx), W8_C255);
  const uint32_t n3 = __vimin3_s16x2_relu(__vadd2(o[5], nx), __vadd2(o[1], ny), W8_C255);
  const uint32_t n4 = __vimin3_s16x2_relu(__vadd2(o[6], nx), __vadd2(o[2], ny), W8_C255);
  const uint32_t n5 = __vimin3_s16x2_relu(__vadd2(o[6], ny), __vadd2(o[2], nx), W8_C255);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}
// branch sums of a forward step, turbodecoder_win.h:724-741: z = information bit 0 into state i, w = information bit 1
__device__ __forceinline__ void w8_abranches(const uint32_t (&o)[8], uint32_t nx, uint32_t ny, uint32_t (&z)[8], uint32_t (&w)[8])
{
  const uint32_t nxy = w8n_nxy(nx, ny);
  z[0] = o[0]; z[1] = w8n_adds(o[3], ny); z[2] = w8n_adds(o[4], ny); z[3] = o[7];
  z[4] = o[1]; z[5] = w8n_adds(o[2], ny); z[6] = w8n_adds(o[5], ny); z[7] = o[6];
  w[0] = w8n_adds(o[1], nxy); w[1] = w8n_adds(o[2], nx); w[2] = w8n_adds(o[5], nx); w[3] = w8n_adds(o[6], nxy);
  w[4] = w8n_adds(o[0], nxy); w[5] = w8n_adds(o[3], nx); w[6] = w8n_adds(o[4], nx); w[7] = w8n_adds(o[7], nxy);
}
// forward step without output (warm-up)
__device__ __forceinline__ void w8_astep(uint32_t (&o)[8], uint32_t nx, uint32_t ny)
{
  const uint32_t nxy = w8n_nxy(nx, ny);
  const uint32_t n0 = w8n_addmin(o[1], nxy, o[0]);
  const uint32_t n3 = w8n_addmin(o[6], nxy, o[7]);
  const uint32_t n4 = w8n_addmin(o[0], nxy, o[1]);
  const uint32_t n7 = w8n_addmin(o[7], nxy, o[6]);
  const uint32_t n1 = __vimin3_s16x2_relu(__vadd2(o[3], ny), __vadd2(o[2], nx), W8_C255);
  const uint32_t n2 = __vimin3_s16x2_relu(__vadd2(o[4], ny), __vadd2(o[5], nx), W8_C255);
  const uint32_t n5 = __vimin3_s16x2_relu(__vadd2(o[2], ny), __vadd2(o[3], nx), W8_C255);
  const uint32_t n6 = __vimin3_s16x2_relu(__vadd2(o[5], ny), __vadd2(o[4], nx), W8_C255);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

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
  // zero-extending: the stored bytes are 0..255 (negated and biased metrics)
  o[0] = __byte_perm(v.x, 0u, 0x4140u); o[1] = __byte_perm(v.x, 0u, 0x4342u);
  o[2] = __byte_perm(v.y, 0u, 0x4140u); o[3] = __byte_perm(v.y, 0u, 0x4342u);
  o[4] = __byte_perm(v.z, 0u, 0x4140u); o[5] = __byte_perm(v.z, 0u, 0x4342u);
  o[6] = __byte_perm(v.w, 0u, 0x4140u); o[7] = __byte_perm(v.w, 0u, 0x4342u);
}
// a load the compiler must leave where it is written (it sinks plain loads to their first use, which puts the whole
// global-memory latency in front of the consumer: measured as a quarter of the stall samples before)
__device__ __forceinline__ uint4 w8_ld_v4(const uint4* p)
{
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// per block size, everything indexed by the window-interleaved position k * NW + w
struct W8Tables {
  const uint16_t* dst1;     // DEC1: where the extrinsic value of position (w, k) goes in app2: (k' * 32 + w') of natural position rev[wS + k]
  const uint16_t* dst2;     // DEC2: the same into app1 / nd through fwd[]
  const uint32_t* crc1[3];  // [crc kind]: x^(K-1-n+24) mod g for the natural position n = wS + k (DEC1 decides in natural order)
  const uint32_t* crc2[3];  // the same for position fwd[wS + k] (DEC2 decides in interleaved order); kind 0 aliases kind 2
};

/*
 * MODE 0: DEC1, first half-iteration     x = syst                         y = par0
 * MODE 1: DEC1 with a-priori             a = app1 - ext1 (glue), x = a + syst   y = par0
 * MODE 2: DEC2                           x = app2                         y = par1
 * grid = n_units, block = 32. ckg: W8_CK_UNIT uint4 per unit.
 */
template <int MODE>
__global__ void __launch_bounds__(32) win8_kernel(const Unit8* __restrict__ units, const W8Tables* __restrict__ tabs, uint8_t* __restrict__ ws,
                                                  uint4* __restrict__ ckg, uint8_t* __restrict__ done, uint8_t* __restrict__ noi,
                                                  uint8_t* __restrict__ ok, uint32_t cnt, uint32_t max_iter, uint32_t min_iter, int early_stop,
                                                  const uint8_t* __restrict__ max_iter_cb)
{
  extern __shared__ __align__(128) uint8_t smem_raw[];
  W8Smem&        sm   = *reinterpret_cast<W8Smem*>(smem_raw);
  const Unit8*   up   = units + blockIdx.x;
  const int      lane = threadIdx.x;
  const uint32_t K = up->K, NW = up->NW, S = up->S, crc_kind = up->crc_kind;
  const int      pair = lane / (int)NW, w_lane = lane % (int)NW;
  // the two code blocks of this lane's pair (pair 1 exists with 16 windows only)
  const int      cb_lo = pair ? up->cb[2] : up->cb[0], cb_hi = pair ? up->cb[3] : up->cb[1];
  const bool     live_lo = cb_lo >= 0 && !done[cb_lo], live_hi = cb_hi >= 0 && !done[cb_hi];
  if (!__any_sync(0xffffffffu, live_lo || live_hi)) return;
  const W8Tables tb = tabs[up->kidx];
  uint8_t*       base = ws + up->ws_off;
  const uint64_t ab   = w8_arr_bytes(S);
  uint4*         ck   = ckg + (size_t)blockIdx.x * W8_CK_UNIT;
  // sub_glue: srsran_vec_sub_bbb saturates, except (AVX2 build) on the last K % 32 elements of the window-interleaved array, which
  // its scalar tail subtracts with wrap-around: interleaved index = step * NW + window >= K - K % 32  <=>  (K % 32 != 0 and step == S-1)
  const bool wrap_tail = (K & 31u) != 0;
  auto glue = [&](uint32_t a, uint32_t b, int k) { return (wrap_tail && k == (int)S - 1) ? w8_wrap8(__vsub2(a, b)) : w8_subs(a, b); };

  // ---------------------------------------------------------------- stage: the x and y streams of the unit -> shared memory
  // 128-bit loads, four trips in flight (every load of a batch is issued before the first value is used)
  {
    const uint4* gx  = reinterpret_cast<const uint4*>(base + (MODE == 2 ? W8_APP2 : W8_SYST) * ab);
    const uint4* gy  = reinterpret_cast<const uint4*>(base + (MODE == 2 ? W8_PAR1 : W8_PAR0) * ab);
    uint4*       ga  = reinterpret_cast<uint4*>(base + W8_APP1 * ab);
    const uint4* ge  = reinterpret_cast<const uint4*>(base + W8_EXT1 * ab);
    uint4*       sx  = reinterpret_cast<uint4*>(sm.X);
    uint4*       sy  = reinterpret_cast<uint4*>(sm.Y);
    const int    nv  = (int)S * (W8_ROW / 16);  // uint4 per stream: four per step row
    constexpr int UN = 4;
    for (int i0 = 0; i0 < nv; i0 += 32 * UN) {
      uint4 vx[UN], vy[UN], va[UN], ve[UN];
#pragma unroll
      for (int t = 0; t < UN; t++) {
        const int i = i0 + 32 * t + lane;
        if (i < nv) {
          vx[t] = gx[i];
          vy[t] = gy[i];
          if (MODE == 1) {
            va[t] = ga[i];
            ve[t] = ge[i];
          }
        }
      }
#pragma unroll
      for (int t = 0; t < UN; t++) {
        const int i = i0 + 32 * t + lane;
        if (i >= nv) continue;
        if (MODE == 1) {
          // app1 <- app1 - ext1 (turbodecoder_iter.h:106-108), x = sat(app1 + syst) (simd_add(ap, x), turbodecoder_win.h:608-611);
          // eight byte pairs per uint4, two per word
          const int      k = i / (W8_ROW / 16);
          const uint32_t xa[4] = {vx[t].x, vx[t].y, vx[t].z, vx[t].w}, aa[4] = {va[t].x, va[t].y, va[t].z, va[t].w},
                         ea[4] = {ve[t].x, ve[t].y, ve[t].z, ve[t].w};
          uint32_t xo[4], ao[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t ap0 = glue(w8_unpack(aa[q]), w8_unpack(ea[q]), k), ap1 = glue(w8_unpack_hi(aa[q]), w8_unpack_hi(ea[q]), k);
            const uint32_t x0 = w8_adds(ap0, w8_unpack(xa[q])), x1 = w8_adds(ap1, w8_unpack_hi(xa[q]));
            ao[q] = __byte_perm(ap0, ap1, 0x6420u);
            xo[q] = __byte_perm(x0, x1, 0x6420u);
          }
          sx[i] = make_uint4(~xo[0], ~xo[1], ~xo[2], ~xo[3]);
          ga[i] = make_uint4(ao[0], ao[1], ao[2], ao[3]);  // the output side subtracts the updated a-priori values again
        } else {
          sx[i] = make_uint4(~vx[t].x, ~vx[t].y, ~vx[t].z, ~vx[t].w);
        }
        sy[i] = make_uint4(~vy[t].x, ~vy[t].y, ~vy[t].z, ~vy[t].w);
      }
    }
  }
  __syncwarp();
  // shared memory holds the complemented bytes: sign-extended ~v = -v - 1, plus one = the NEGATED branch metric
  auto ldx = [&](int k) { return __vadd2(w8_unpack(*reinterpret_cast<const uint16_t*>(&sm.X[k * W8_ROW + 2 * lane])), W8_ONE); };
  auto ldy = [&](int k) { return __vadd2(w8_unpack(*reinterpret_cast<const uint16_t*>(&sm.Y[k * W8_ROW + 2 * lane])), W8_ONE); };

  uint32_t o[8];
  // ---------------------------------------------------------------- beta: warm-up over the window's own first 40 steps
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = W8_P127;  // simd_set1(-INF), INF = 0; negated and biased: 127
#pragma unroll 2
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
      const int8_t* tail = reinterpret_cast<const int8_t*>(base + W8_NARR * ab);  // [cb slot][stream: syst, par0, app2, par1][4]
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int st0 = 0, st1 = 0, st2 = 0, st3 = 0, st4 = 0, st5 = 0, st6 = 0, st7 = 0;
        if ((h ? cb_hi : cb_lo) >= 0) {
          const int8_t* pin = tail + (2 * pair + h) * 16 + (MODE == 2 ? 8 : 0);   // app2 (second encoder's termination systematic) / syst
          const int8_t* ppa = tail + (2 * pair + h) * 16 + (MODE == 2 ? 12 : 4);  // par1 / par0
          for (int r = 2; r >= 0; r--) {
            const int xv = pin[r], yv = ppa[r], xy = w8_sadd_tail(xv, yv);
            const int n0 = max(w8_sadd_tail(st4, xy), st0), n1 = max(st4, w8_sadd_tail(st0, xy));
            const int n2 = max(w8_sadd_tail(st5, yv), w8_sadd_tail(st1, xv)), n3 = max(w8_sadd_tail(st5, xv), w8_sadd_tail(st1, yv));
            const int n4 = max(w8_sadd_tail(st6, xv), w8_sadd_tail(st2, yv)), n5 = max(w8_sadd_tail(st6, yv), w8_sadd_tail(st2, xv));
            const int n6 = max(st7, w8_sadd_tail(st3, xy)), n7 = max(w8_sadd_tail(st7, xy), st3);
            st0 = n0; st1 = n1; st2 = n2; st3 = n3; st4 = n4; st5 = n5; st6 = n6; st7 = n7;
          }
        }
        const int sv[8] = {st0, st1, st2, st3, st4, st5, st6, st7};
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = h ? ((t[i] & 0xffffu) | ((uint32_t)(127 - sv[i]) << 16)) : (uint32_t)(127 - sv[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = t[i];
  }
  // ---------------------------------------------------------------- beta: main pass, checkpoint B[k] (before normalisation) at k = 8, 16, ... and B[S]
  const int ck_top = ((int)S + 7) / 8;  // slot of B[S]; B[8c] lives in slot c
  w8_ck_store(&ck[ck_top * 32 + lane], o);
#pragma unroll 2
  for (int k = (int)S - 1; k >= 0; k--) {
    w8_bstep(o, ldx(k), ldy(k));
    if (k && (k & 7) == 0) w8_ck_store(&ck[(k >> 3) * 32 + lane], o);
    if (k) w8_norm(o);
  }
  // ---------------------------------------------------------------- alpha: warm-up over the window's own last 40 steps
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = W8_P127;
#pragma unroll 2
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
    for (int i = 0; i < 8; i++) o[i] = (w_lane == 0) ? W8_P127 : t[i];  // first window: the known state {0, -INF x 7}, INF = 0
  }
  // ---------------------------------------------------------------- alpha: main pass, 8 steps at a time with beta recomputed into registers
  // Every step also finishes its own output: the extrinsic value goes to ext1 in place (DEC1) and through the QPP table to the
  // other decoder's input, the hard decision to its natural position, the CRC contribution into the running remainder. All of
  // these are 2-byte accesses of one 64-byte row per warp - or one byte when the other block of the pair is done and must keep
  // what it has.
  const uint16_t* dtab  = (MODE == 2) ? tb.dst2 : tb.dst1;
  const uint32_t* ctab  = (MODE == 2) ? tb.crc2[crc_kind] : tb.crc1[crc_kind];
  uint8_t*        pOUT  = base + (MODE == 2 ? W8_APP1 : W8_APP2) * ab;  // scatter target
  uint8_t*        pE1   = base + W8_EXT1 * ab;
  uint8_t*        pND   = base + W8_ND * ab;
  const uint8_t*  pAP   = base + W8_APP1 * ab;
  const uint32_t  lane2 = 2u * (uint32_t)lane, pcol = 2u * (uint32_t)(pair * (int)NW);
  uint32_t        crc_lo = 0, crc_hi = 0;
  // a 2-byte pair store that leaves the byte of a finished code block alone
  auto st_pair = [&](uint8_t* p, uint32_t v) {
    if (live_lo && live_hi) *reinterpret_cast<uint16_t*>(p) = (uint16_t)v;
    else if (live_lo) p[0] = (uint8_t)v;
    else if (live_hi) p[1] = (uint8_t)(v >> 8);
  };
  uint4 ck_next = w8_ld_v4(&ck[(8 >= (int)S ? ck_top : 1) * 32 + lane]);
  for (int k0 = 0; k0 < (int)S; k0 += 8) {
    const int len = min(8, (int)S - k0), top = k0 + len;
    // per-step table entries and a-priori values of this block, fetched before the block's ~900 instructions
    uint32_t dv[8], cv[8], av[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int kk = min(k0 + j, (int)S - 1);
      dv[j] = __ldg(dtab + kk * (int)NW + w_lane);
      cv[j] = __ldg(ctab + kk * (int)NW + w_lane);
      av[j] = (MODE == 1) ? (uint32_t) * reinterpret_cast<const uint16_t*>(pAP + kk * W8_ROW + lane2) : 0u;
    }
    // B[j] = beta[k0 + 1 + j] as stored (before normalisation): what the LLR of step k0 + j reads; B[len-1] is the checkpoint
    uint32_t s[8];
    w8_ck_unpack(ck_next, s);
    if (top < (int)S) {  // the next block's checkpoint is fetched now and used after this block
      const int ntop = min(top + 8, (int)S);
      ck_next        = w8_ld_v4(&ck[(ntop == (int)S ? ck_top : (ntop >> 3)) * 32 + lane]);
    }
    // one step: branch sums, LLR against beta[k+1] (b), outputs, state update
    auto astep_out = [&](int k, uint32_t dvj, uint32_t cvj, uint32_t avj, const uint32_t (&b)[8]) {
      const uint32_t nx = ldx(k), ny = ldy(k);
      uint32_t       z[8], w[8];
      w8_abranches(o, nx, ny, z, w);
      // max_i sat(b_i + z_i) = sat(max_i (b_i + z_i)): saturation is monotone. Negated: b" + z" = 254 - (b + z), the maximum is
      // the minimum M of those, and 127 - sat(254 - M) = max(min(M - 127, 255), 0)
      const uint32_t p0 = __vadd2(b[0], z[0]), p1 = __vadd2(b[1], z[1]), p2 = __vadd2(b[2], z[2]), p3 = __vadd2(b[3], z[3]);
      const uint32_t p4 = __vadd2(b[4], z[4]), p5 = __vadd2(b[5], z[5]), p6 = __vadd2(b[6], z[6]), p7 = __vadd2(b[7], z[7]);
      const uint32_t q0 = __vadd2(b[0], w[0]), q1 = __vadd2(b[1], w[1]), q2 = __vadd2(b[2], w[2]), q3 = __vadd2(b[3], w[3]);
      const uint32_t q4 = __vadd2(b[4], w[4]), q5 = __vadd2(b[5], w[5]), q6 = __vadd2(b[6], w[6]), q7 = __vadd2(b[7], w[7]);
      const uint32_t m0 = __viaddmin_s16x2_relu(__vmins2(__vimin3_s16x2(__vimin3_s16x2(p0, p1, p2), p6, p7), __vimin3_s16x2(p3, p4, p5)), W8_M127, W8_C255);
      const uint32_t m1 = __viaddmin_s16x2_relu(__vmins2(__vimin3_s16x2(__vimin3_s16x2(q0, q1, q2), q6, q7), __vimin3_s16x2(q3, q4, q5)), W8_M127, W8_C255);
      const uint32_t l  = w8_subs(m0, m1);  // sat(max1 - max0) = sat(m0" - m1")
      // out = l >> 1, arithmetic, per element (simd_rb_shift, divide_output = 1)
      const uint32_t ext = ((l >> 1) & 0x7fff7fffu) | (l & 0x80008000u);
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = __vmins2(z[i], w[i]);
      if (k) w8_norm(o);
      // what the next half-iteration reads: DEC1 ext1 <- ext1 - app1 (turbodecoder_iter.h:116-118; first half-iteration: ext1
      // as it is), app2[rev[i]] = ext1[i] (:120); DEC2 app1[fwd[i]] = ext2[i] (:127)
      uint32_t r = ext;
      if (MODE == 1) r = glue(ext, w8_unpack(avj), k);
      const uint32_t rp = w8_pack(r);
      if (MODE != 2) st_pair(pE1 + k * W8_ROW + lane2, rp);
      st_pair(pOUT + 2u * dvj + pcol, rp);
      // the decision (tdec_win*_decision_byte: ext > 0, i.e. l >= 2) at its natural position - stored as 0xFF / 0x00 - and its CRC
      // contribution: l + 0x7FFE has its sign bit set exactly when l >= 2; PRMT replicates the sign of a byte over a byte
      const uint32_t sg = __vadd2(l, 0x7ffe7ffeu);
      st_pair(pND + (MODE == 2 ? 2u * dvj + pcol : (uint32_t)k * W8_ROW + lane2), w8_prmt(sg, 0u, 0x44B9u));
      crc_lo ^= cvj & w8_prmt(sg, 0u, 0x9999u);
      crc_hi ^= cvj & w8_prmt(sg, 0u, 0xBBBBu);
    };
    if (len == 8) {
      uint32_t B[8][8];
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
      for (int j = 0; j < 8; j++) astep_out(k0 + j, dv[j], cv[j], av[j], B[j]);
    } else {
      // the short last block of a window whose length is not a multiple of 8 (top == S, at most 7 steps, once per window): beta of
      // every step is recomputed from the checkpoint on its own - no register window with a runtime length (an array indexed
      // by `len` would push the whole window into local memory, for the full blocks too)
#pragma unroll
      for (int j = 0; j < 7; j++) {
        if (j < len) {
          uint32_t t[8];
#pragma unroll
          for (int i = 0; i < 8; i++) t[i] = s[i];
          if (j < len - 1 && top < (int)S) w8_norm(t);
          for (int kk = top - 1; kk >= k0 + 1 + j; kk--) {
            w8_bstep(t, ldx(kk), ldy(kk));
            if (kk > k0 + 1 + j) w8_norm(t);  // kk >= 1; the last one stays as stored: before normalisation
          }
          astep_out(k0 + j, dv[j], cv[j], av[j], t);
        }
      }
    }
  }
  // ---------------------------------------------------------------- verdict per code block (sch.c:426-456): CRC of all windows, counts, done flags
  for (int off = (int)NW / 2; off > 0; off >>= 1) {
    crc_lo ^= __shfl_xor_sync(0xffffffffu, crc_lo, off, (int)NW);
    crc_hi ^= __shfl_xor_sync(0xffffffffu, crc_hi, off, (int)NW);
  }
  if (w_lane == 0) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int cb = h ? cb_hi : cb_lo;
      if (cb < 0 || !(h ? live_hi : live_lo)) continue;
      const uint32_t okv = (crc_kind != 0 && (h ? crc_hi : crc_lo) == 0u) ? 1u : 0u;
      noi[cb] = (uint8_t)cnt;
      ok[cb]  = (uint8_t)okv;
      if ((early_stop && okv && cnt >= min_iter) || cnt >= (max_iter_cb ? (uint32_t)max_iter_cb[cb] : max_iter)) done[cb] = 1;
    }
  }
}

/*
 * De-multiplex natural-order int8 LLRs (s p p' triples + 12 termination values: tdec_win*_extract_input, turbodecoder_win.h:
 * 883-921) into the unit's window-interleaved, pair-packed arrays and re-arm the decode state. One block per unit: the triples
 * are read coalesced, transposed in shared memory and leave as whole 64-byte rows. Dynamic shared memory: 3 * S * 66 bytes.
 */
__global__ void __launch_bounds__(256) extract8_kernel(const int8_t* __restrict__ llr, const uint64_t* __restrict__ llr_off, const Unit8* __restrict__ units,
                                                       uint8_t* __restrict__ ws, uint8_t* __restrict__ done)
{
  extern __shared__ __align__(16) uint8_t xs[];  // [3][S][66]: rows padded by 2 bytes so that the transposing stores spread over the banks
  const Unit8    u = units[blockIdx.x];
  const uint32_t K = u.K, NW = u.NW, S = u.S, PITCH = W8_ROW + 2;
  uint8_t*       base = ws + u.ws_off;
  const uint64_t ab   = w8_arr_bytes(S);
  const int      ncb  = 2 * (32 / (int)NW);
  for (uint32_t i = threadIdx.x; i < 3 * S * PITCH / 2; i += 256) reinterpret_cast<uint16_t*>(xs)[i] = 0;
  __syncthreads();
  for (int c = 0; c < ncb; c++) {
    const int cb = u.cb[c];
    if (cb < 0) continue;
    const int8_t*  in  = llr + llr_off[cb];
    const uint32_t col = 2u * ((c >> 1) * NW) + (c & 1);
    for (uint32_t n = threadIdx.x; n < K; n += 256) {
      const uint32_t w = n / S, k = n - w * S, a = k * PITCH + 2 * w + col;
      xs[a]                 = (uint8_t)in[3 * n];
      xs[S * PITCH + a]     = (uint8_t)in[3 * n + 1];
      xs[2 * S * PITCH + a] = (uint8_t)in[3 * n + 2];
    }
    if (threadIdx.x < 3) {
      const uint32_t j = threadIdx.x;
      int8_t* tail = reinterpret_cast<int8_t*>(base + W8_NARR * ab) + c * 16;
      tail[j]      = in[3 * K + 2 * j];          // syst termination
      tail[4 + j]  = in[3 * K + 2 * j + 1];      // par0
      tail[8 + j]  = in[3 * K + 6 + 2 * j];      // app2: second encoder's termination systematic
      tail[12 + j] = in[3 * K + 6 + 2 * j + 1];  // par1
    }
    if (threadIdx.x == 0) done[cb] = 0;
  }
  __syncthreads();
  // rows out: 3 streams x S rows x 16 words
  for (uint32_t i = threadIdx.x; i < 3 * S * 16; i += 256) {
    const uint32_t a = i / (S * 16), r = i - a * S * 16, k = r / 16, wq = r - k * 16;
    const uint8_t* src = xs + a * S * PITCH + k * PITCH + 4 * wq;
    const uint32_t v   = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    reinterpret_cast<uint32_t*>(base + a * ab)[k * 16 + wq] = v;  // arrays 0, 1, 2 = syst, par0, par1
  }
}

/*
 * Decoded bytes of every code block from the unit's decision array (tdec_win*_decision_byte, MSB first): position n = w S + k sits
 * in row k, column w of its pair. One warp per unit; 32 positions per ballot.
 */
__global__ void __launch_bounds__(32) emit8_kernel(const Unit8* __restrict__ units, const uint8_t* __restrict__ ws, uint8_t* __restrict__ out,
                                                   const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_len)
{
  const Unit8    u    = units[blockIdx.x];
  const int      lane = threadIdx.x;
  const uint32_t K = u.K, NW = u.NW, S = u.S;
  const uint8_t* nd  = ws + u.ws_off + W8_ND * w8_arr_bytes(S);
  const int      ncb = 2 * (32 / (int)NW);
  for (int c = 0; c < ncb; c++) {
    const int cb = u.cb[c];
    if (cb < 0) continue;
    const uint32_t col   = 2u * ((c >> 1) * NW) + (c & 1);
    uint8_t*       dst   = out + out_off[cb];
    const uint32_t total = out_len ? out_len[cb] : K / 8;
    uint32_t       w = 0, k = (uint32_t)lane;  // position n0 + lane walks (window, step) incrementally: S > 32
    for (uint32_t n0 = 0; n0 < K; n0 += 32) {
      const uint32_t n = n0 + lane;
      const uint32_t d = (n < K) ? nd[k * W8_ROW + 2 * w + col] : 0u;
      k += 32;
      if (k >= S) { k -= S; w++; }
      const uint32_t word = __brev(__ballot_sync(0xffffffffu, d != 0));  // position n0 in bit 31: MSB-first bytes
      const uint32_t b    = (n0 >> 3) + (uint32_t)lane;
      if (lane < 4 && b < total && b < K / 8) dst[b] = (uint8_t)(word >> (24 - 8 * lane));
    }
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
constexpr uint32_t RM8_SMEM    = 32768;  // e-bits staged per block; the rest of a longer e (heavy repetition) is read from global
constexpr int      RM8_THREADS = 512;
// grid = code blocks, block = RM8_THREADS, dynamic shared memory = min(max E, RM8_SMEM) + 16 bytes
__global__ void __launch_bounds__(RM8_THREADS) rm_rx8_kernel(const RmJob8* __restrict__ jobs)
{
  extern __shared__ __align__(16) int8_t se8_raw[];
  const RmJob8   j  = jobs[blockIdx.x];
  const uint32_t ns = min(j.E, RM8_SMEM);
  // staged copy shifted like the source, so that 128-bit loads and stores line up (see rm_rx_kernel)
  const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(j.e) & 15u);
  int8_t*        se = se8_raw + sh;
  {
    const uint32_t head = min(ns, (16u - sh) & 15u);
    if (threadIdx.x < head) se[threadIdx.x] = j.e[threadIdx.x];
    const uint32_t nv = (ns - head) / 16;
    const uint4*   gv = reinterpret_cast<const uint4*>(j.e + head);
    uint4*         sv = reinterpret_cast<uint4*>(se + head);
    for (uint32_t i = threadIdx.x; i < nv; i += RM8_THREADS) sv[i] = __ldg(gv + i);
    for (uint32_t i = head + 16 * nv + threadIdx.x; i < ns; i += RM8_THREADS) se[i] = j.e[i];
  }
  __syncthreads();
  const uint32_t reps = (j.E + j.L - 1) / j.L;  // block-uniform trip count over the repetitions (1 without repetition)
  const uint32_t se_addr = (uint32_t)__cvta_generic_to_shared(se);  // (see rm_rx_kernel: one address computation, not one per load)
  auto llr = [&](uint32_t i) -> uint32_t {
    if (i >= ns) return (uint8_t)j.e[i];
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(se_addr + i));
    return v;
  };
  // four soft-buffer positions (one 32-bit word, L = 3K + 12 is a multiple of 4) per thread: 64-bit table load, wrapping byte adds
  const bool     wide = (reinterpret_cast<uintptr_t>(j.buf) & 3u) == 0 && (reinterpret_cast<uintptr_t>(j.table) & 7u) == 0;
  const uint32_t L4   = wide ? j.L / 4 : 0u;
  for (uint32_t p4 = threadIdx.x; p4 < L4; p4 += RM8_THREADS) {
    const uint2    tt   = __ldg(reinterpret_cast<const uint2*>(j.table) + p4);
    const uint32_t n[4] = {tt.x & 0xffffu, tt.x >> 16, tt.y & 0xffffu, tt.y >> 16};
    if (n[0] >= j.E && n[1] >= j.E && n[2] >= j.E && n[3] >= j.E) continue;  // punctured region
    const uint32_t old    = reinterpret_cast<uint32_t*>(j.buf)[p4];
    uint32_t       sum[4] = {0u, 0u, 0u, 0u};
    for (uint32_t r = 0, base = 0; r < reps; r++, base += j.L) {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (n[q] + base < j.E) sum[q] += llr(n[q] + base);
    }
    const uint32_t add = (sum[0] & 0xffu) | ((sum[1] & 0xffu) << 8) | ((sum[2] & 0xffu) << 16) | (sum[3] << 24);
    reinterpret_cast<uint32_t*>(j.buf)[p4] = __vadd4(old, add);
  }
  for (uint32_t p = 4 * L4 + threadIdx.x; p < j.L; p += RM8_THREADS) {
    const uint32_t t = j.table[p];
    if (t >= j.E) continue;
    uint32_t acc = 0;
    for (uint32_t i = t; i < j.E; i += j.L) acc += llr(i);
    j.buf[p] = (int8_t)(uint8_t)((uint8_t)j.buf[p] + acc);
  }
}

}  // namespace srsb200
