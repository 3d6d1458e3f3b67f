/*
 * rm_kernels.cuh - rate de-matching + HARQ soft combining (srsran_rm_turbo_rx_lut, lib/src/phy/fec/turbo/rm_turbo.c:390-445),
 * transport-block CRC24A (decode_tb, lib/src/phy/phch/sch.c:563), the UL-SCH channel de-interleaver and the gather copy as
 * sm_100a kernels.
 *
 * The reference computes  output[T[i mod L]] += input[i]  for i < E  (L = 3K+12, int16 wrap). T is a permutation of
 * 0..L-1, so soft-buffer position p receives exactly the received LLRs  e[n], e[n+L], e[n+2L], ...  with n = Tinv[p]
 * (repetition when E > L). The kernel runs in that GATHER form: a block owns a code block, stages its e-bits in shared
 * memory (13 KB at the 100-PRB 64QAM operating point) and walks the soft buffer in order - the read-modify-write of the
 * 35 KB soft buffer is coalesced, the permuted reads hit shared memory. (The first version scattered: thread n did
 * buf[T[n]] += e[n]; the sub-block interleaver puts consecutive n 192 bytes apart, so every 2-byte read-modify-write
 * cost two 32-byte sectors - 338 GB/s, 5 % of the HBM peak, measured on 6656 code blocks.) Bit-identical to the
 * reference because 16-bit wrapping addition is associative and commutative.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace srsb200 {

struct RmJob {
  const int16_t*  e;      // first received LLR of this code block (device)
  int16_t*        buf;    // soft buffer of this code block, natural layout (device, 4-byte aligned)
  const uint16_t* table;  // Tinv[0..L): position in the transmitted (circular-buffer) order of soft-buffer element p
  uint32_t        E;
  uint32_t        L;
  // optional descrambling (36.211 7.2) fused into the gather: the LLRs are still scrambled, e[i] belongs to position
  // c_off + i of the codeword's scrambling sequence; x1 / x2 = LFSR windows at sequence position 0 (i.e. after Nc = 1600)
  uint32_t        scramble, c_off, x1, x2;
};

constexpr uint32_t RM_SMEM_ELEMS = 24576;  // e-bits staged per block (48 KB); the rest of a longer e is read from global
constexpr int      RM_THREADS    = 512;
constexpr uint32_t RM_SC_WORDS   = 1024;   // scrambling bits held per block (32768 = more than any code block receives unrepeated)
constexpr int      GOLD_POWERS   = 20;     // jump matrices A^(2^k), k < 20: offsets up to 2^20 sequence positions

// y = M x over GF(2) for a 31-bit LFSR window; M as 31 row masks (32 words, the last unused)
__device__ __forceinline__ uint32_t gold_matvec(const uint32_t* __restrict__ rows, uint32_t x)
{
  uint32_t y = 0;
#pragma unroll
  for (int r = 0; r < 31; r++) y |= (uint32_t)(__popc(__ldg(rows + r) & x) & 1) << r;
  return y;
}
// window of an LFSR `off` positions later: jump[k] = A^(2^k) (32 words each)
__device__ __forceinline__ uint32_t gold_jump(const uint32_t* __restrict__ jump, uint32_t x, uint32_t off)
{
  for (int k = 0; off; k++, off >>= 1)
    if (off & 1u) x = gold_matvec(jump + 32 * k, x);
  return x;
}

// the same products with the 31 rows spread over the lanes of a warp (x and off are warp-uniform)
__device__ __forceinline__ uint32_t warp_gold_matvec(const uint32_t* __restrict__ rows, uint32_t x, int lane)
{
  const uint32_t bit = lane < 31 ? (uint32_t)(__popc(__ldg(rows + lane) & x) & 1) : 0u;
  return __ballot_sync(0xffffffffu, bit);
}
__device__ __forceinline__ uint32_t warp_gold_jump(const uint32_t* __restrict__ jump, uint32_t x, uint32_t off, int lane)
{
  for (int k = 0; off; k++, off >>= 1)
    if (off & 1u) x = warp_gold_matvec(jump + 32 * k, x, lane);
  return x;
}

// grid = n_jobs, block = RM_THREADS, dynamic shared memory = min(max E, RM_SMEM_ELEMS) * 2 bytes.
// gold = jump matrices of the two LFSRs ([2][GOLD_POWERS][32] words), only read by jobs with scramble != 0
// SCR = false: no job of the launch descrambles (the sequence staging and the sign test are compiled out)
template <bool SCR>
__global__ void __launch_bounds__(RM_THREADS, 4) rm_rx_kernel(const RmJob* __restrict__ jobs, const uint32_t* __restrict__ gold)
{
  extern __shared__ __align__(16) int16_t se_raw[];
  __shared__ uint32_t sc[SCR ? RM_SC_WORDS : 1];
  const RmJob    j  = jobs[blockIdx.x];
  const uint32_t nthr = blockDim.x;  // <= RM_THREADS
  const uint32_t ns = min(j.E, RM_SMEM_ELEMS);
  // ---- stage the received LLRs: 128-bit global loads and 128-bit shared stores. The e-bits of a code block start anywhere
  //      (2-byte aligned): the shared copy is shifted by the same number of elements, so that element i sits 16-byte aligned in
  //      shared memory exactly when it does in global memory; a scalar head and tail cover the rest.
  const uint32_t sh = (uint32_t)((reinterpret_cast<uintptr_t>(j.e) >> 1) & 7u);  // elements past a 16-byte boundary
  int16_t*       se = se_raw + sh;                                               // (8 spare elements are allocated for the shift)
  {
    const uint32_t head = min(ns, (8u - sh) & 7u);
    if (threadIdx.x < head) se[threadIdx.x] = j.e[threadIdx.x];
    const uint32_t nv = (ns - head) / 8;
    const uint4*   gv = reinterpret_cast<const uint4*>(j.e + head);
    uint4*         sv = reinterpret_cast<uint4*>(se + head);
    for (uint32_t i = threadIdx.x; i < nv; i += nthr) sv[i] = __ldg(gv + i);
    for (uint32_t i = head + 8 * nv + threadIdx.x; i < ns; i += nthr) se[i] = j.e[i];
  }
  if (SCR && j.scramble) {
    // The block's part of the scrambling sequence, 32 bits per word. A matrix-vector product over GF(2) is one popc per row:
    // the 32 lanes of a warp take one row each and a ballot collects the new window, so a warp jumps both LFSR windows to
    // its first word (offset c_off + 32 * first word, by powers A^(2^k)) and then walks word by word with A^32. 31 bits of a
    // word come straight out of the windows (bit b = x(n + b)), the 32nd is the feedback bit.
    const int      lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t nw   = min((j.E + 31) / 32, RM_SC_WORDS), per = (nw + nthr / 32 - 1) / (nthr / 32);
    const uint32_t w0 = wid * per, w1 = min(w0 + per, nw);
    if (w0 < w1) {
      uint32_t a = warp_gold_jump(gold, j.x1, j.c_off + 32 * w0, lane), b = warp_gold_jump(gold + 32 * GOLD_POWERS, j.x2, j.c_off + 32 * w0, lane);
      for (uint32_t w = w0; w < w1; w++) {
        const uint32_t f = ((a ^ (a >> 3)) ^ (b ^ (b >> 1) ^ (b >> 2) ^ (b >> 3))) & 1u;
        if (lane == 0) sc[w] = ((a ^ b) & 0x7fffffffu) | (f << 31);
        a = warp_gold_matvec(gold + 32 * 5, a, lane);                      // A^32
        b = warp_gold_matvec(gold + 32 * GOLD_POWERS + 32 * 5, b, lane);
      }
    }
  }
  __syncthreads();
  // The staged copy is read through its 32-bit shared-memory address, computed once: through the shifted generic pointer the
  // compiler rebuilt the shared window base (S2R + MOV + LEA) in front of every predicated load - 20 instructions per LLR, 6 now.
  const uint32_t se_addr = (uint32_t)__cvta_generic_to_shared(se);
  auto lds16 = [&](uint32_t i) -> uint32_t {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(se_addr + 2u * i));
    return v;
  };
  // one received LLR (descrambled if asked for). STAGED: every i < E lies in shared memory (E <= RM_SMEM_ELEMS, the usual case)
  auto llr = [&](uint32_t i, auto staged) -> uint32_t {
    uint32_t v = (decltype(staged)::value || i < ns) ? lds16(i) : (uint32_t)(uint16_t)j.e[i];
    if (SCR && j.scramble) {
      uint32_t c;
      if ((i >> 5) < RM_SC_WORDS) {
        c = (sc[i >> 5] >> (i & 31u)) & 1u;
      } else {  // beyond the staged part (heavy repetition): one jump per element, rare
        c = (gold_jump(gold, j.x1, j.c_off + i) ^ gold_jump(gold + 32 * GOLD_POWERS, j.x2, j.c_off + i)) & 1u;
      }
      if (c) v = 0u - v;  // int16 wrap: -(-32768) stays -32768 (sequence.c:545)
    }
    return v;
  };
  // What soft-buffer position p receives: e[n], e[n + L], ... with n = Tinv[p] (repetition when E > L). The trip count over the
  // repetitions is the same for the whole block (reps = ceil(E / L), 1 at every operating point without repetition), so the
  // shared-memory reads of a thread's eight positions are independent, predicated loads issued back to back (the first round-2
  // version walked i = n, n + L, ... per position: eight data-dependent loops, 27 instructions per LLR, issue-bound).
  const uint32_t reps = (j.E + j.L - 1) / j.L;
  // ---- read-modify-write of the soft buffer, EIGHT positions (one uint4) per thread: the table entries and the old values
  //      are fetched with one 128-bit load each before the first shared-memory read, results leave with one 128-bit store.
  //      Table and soft buffer are 16-byte aligned (the engine allocates them so); L = 3K + 12 = 4 mod 8: the last four
  //      positions are done two at a time.
  const bool     wide = ((reinterpret_cast<uintptr_t>(j.buf) | reinterpret_cast<uintptr_t>(j.table)) & 15u) == 0;
  const uint32_t L8   = wide ? j.L / 8 : 0u;
  auto rmw = [&](auto staged) {
    // U uint4 per thread and trip: the table loads and the old values are in flight before the first shared-memory read
    constexpr int U = 1;  // (two per trip - four 128-bit loads in flight per thread - measured slower: 0.163 against 0.127 ms)
    for (uint32_t p0 = threadIdx.x; p0 < L8; p0 += U * nthr) {
      uint4 tt[U], old[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t p8 = p0 + u * nthr;
        if (p8 < L8) {
          tt[u]  = __ldg(reinterpret_cast<const uint4*>(j.table) + p8);
          old[u] = reinterpret_cast<const uint4*>(j.buf)[p8];
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t p8 = p0 + u * nthr;
        if (p8 >= L8) continue;
        const uint32_t tw[4] = {tt[u].x, tt[u].y, tt[u].z, tt[u].w};
        uint32_t       n[8];
        bool           any = false;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          n[2 * q]     = tw[q] & 0xffffu;
          n[2 * q + 1] = tw[q] >> 16;
          any |= n[2 * q] < j.E || n[2 * q + 1] < j.E;
        }
        if (!any) continue;  // nothing received for these eight positions (punctured region): nothing written
        uint32_t ow[4]  = {old[u].x, old[u].y, old[u].z, old[u].w};
        uint32_t sum[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        for (uint32_t r = 0, base = 0; r < reps; r++, base += j.L) {
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint32_t i = n[q] + base;
            if (i < j.E) sum[q] += llr(i, staged);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) ow[q] = __vadd2(ow[q], __byte_perm(sum[2 * q], sum[2 * q + 1], 0x5410));  // two wrapping 16-bit sums
        reinterpret_cast<uint4*>(j.buf)[p8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    // two soft-buffer elements per thread for the rest (L is even, buf is 4-byte aligned)
    uint32_t* buf2 = reinterpret_cast<uint32_t*>(j.buf);
    for (uint32_t p2 = 4 * L8 + threadIdx.x; p2 < j.L / 2; p2 += nthr) {
      const uint32_t tt = reinterpret_cast<const uint32_t*>(j.table)[p2];
      const uint32_t n0 = tt & 0xffffu, n1 = tt >> 16;
      if (n0 >= j.E && n1 >= j.E) continue;  // nothing received for these two positions
      uint32_t s0 = 0, s1 = 0;
      for (uint32_t r = 0, base = 0; r < reps; r++, base += j.L) {
        if (n0 + base < j.E) s0 += llr(n0 + base, staged);
        if (n1 + base < j.E) s1 += llr(n1 + base, staged);
      }
      buf2[p2] = __vadd2(buf2[p2], __byte_perm(s0, s1, 0x5410));
    }
  };
  if (j.E <= RM_SMEM_ELEMS) rmw(std::true_type{}); else rmw(std::false_type{});
}

// zero a list of soft-buffer mirrors (srsran_softbuffer_rx_reset forwarded to the device): grid = (ceil(n16/256), n_slots)
__global__ void __launch_bounds__(256) zero_slots_kernel(int16_t* const* __restrict__ slots, uint32_t n_uint4)
{
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < n_uint4) reinterpret_cast<uint4*>(slots[blockIdx.y])[i] = make_uint4(0u, 0u, 0u, 0u);
}

/*
 * CRC of a byte string by linearity: crc = XOR over set bits of x^(nbits-1-pos+24) mod g. words[m] = x^(m+24) mod g.
 * One block per job; crc_out[job] must be zero on entry (atomicXor accumulation).
 */
struct CrcJob {
  const uint8_t* data;
  uint32_t       nbits;  // multiple of 8
};
__global__ void __launch_bounds__(256) crc_bytes_kernel(const CrcJob* __restrict__ jobs, const uint32_t* __restrict__ words, uint32_t* __restrict__ crc_out)
{
  const CrcJob j   = jobs[blockIdx.y];
  uint32_t     acc = 0;
  // the eight bits of a byte use eight CONSECUTIVE table entries (bit 7 - the earliest - the highest index): two 128-bit
  // loads and eight masked XORs per byte, no data-dependent loop (nbits is a multiple of 8, so the entries are 32-byte aligned)
  const uint32_t nbytes = j.nbits / 8;
  for (uint32_t b0 = (blockIdx.x * 256 + threadIdx.x) * 4; b0 < nbytes; b0 += gridDim.x * 1024) {  // four bytes per thread and trip
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t byte = b0 + k;
      if (byte >= nbytes) break;
      const uint32_t v = j.data[byte];
      if (v == 0) continue;
      const uint4* w  = reinterpret_cast<const uint4*>(words + (j.nbits - 8 - byte * 8));  // entries for bits 0..7 of the byte
      const uint4  lo = __ldg(w), hi = __ldg(w + 1);
      acc ^= (lo.x & (0u - (v & 1u))) ^ (lo.y & (0u - ((v >> 1) & 1u))) ^ (lo.z & (0u - ((v >> 2) & 1u))) ^ (lo.w & (0u - ((v >> 3) & 1u)));
      acc ^= (hi.x & (0u - ((v >> 4) & 1u))) ^ (hi.y & (0u - ((v >> 5) & 1u))) ^ (hi.z & (0u - ((v >> 6) & 1u))) ^ (hi.w & (0u - (v >> 7)));
    }
  }
  acc = __reduce_xor_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(&crc_out[blockIdx.y], acc);
}

/*
 * UL-SCH channel de-interleaver (36.212 5.2.2.8; ulsch_deinterleave, lib/src/phy/phch/sch.c:994-1021). The reference
 * builds lut[] with a running counter over the rows x cols x Qm matrix in (row, column, bit) order - position
 * p = row*Qm + col*rows*Qm + bit - skipping the positions that carry RI (lut = 0 there), then scatters g[lut[p]] = q[p]
 * for ascending p. Here a thread owns one (row, column) pair, i.e. Qm consecutive matrix elements in that order: the rank
 * of an element is its scan index minus the number of RI elements before it (one binary search per pair in the job's
 * sorted list), so writes are coalesced, reads are Qm-value runs, and no table is built. g[0] receives
 * q[p_star], the last position the sequential loop would have written there (an RI position when there is one).
 */
struct DeintJob {
  const int16_t*  q;
  int16_t*        g;
  const uint32_t* ri_scan;  // sorted scan-order indices of the RI positions
  uint32_t        nri, rows, cols, Qm, p_star;
};
// The matrix is rows x cols items of Qm LLRs; q holds it column by column, g row by row. A block takes DT_ROWS consecutive rows:
// per column that is ONE contiguous run of DT_ROWS * Qm LLRs in q - read with coalesced 32-bit loads, several in flight per
// thread, into shared memory - and the block's part of g is one contiguous range. A tile without RI positions inside (all but
// the last rows of a subframe) leaves as a linear stream of 32-bit words, word w of the range taken from (row, column, bit) =
// w's scan position (divisions by multiply-high with per-block constants); tiles with RI positions, or behind an odd number of
// them, go item by item with the rank found by binary search. Round 1 let every thread read its Qm values straight from q
// (12 bytes out of every 32-byte sector per request): 1.5 TB/s; 32-row tiles with per-item stores: 2.1 TB/s.
// grid = jobs * tpj (tpj = blocks per job: any number >= 1, the tiles of a job are strided over them), block = 256
constexpr int DT_ROWS = 128;
constexpr int DT_MAXC = 14;  // columns = PUSCH symbols carrying data: 12 (normal CP), 10 / 11 with SRS or extended CP
__global__ void __launch_bounds__(256) ulsch_deint_kernel(const DeintJob* __restrict__ jobs, uint32_t tpj)
{
  constexpr int  PITCH = DT_ROWS * 8 + 8;  // int16 per tile row: a multiple of 8 (128-bit shared stores), 260 words = 4 mod 32 banks
  __shared__ __align__(16) int16_t tile[DT_MAXC][PITCH];
  // one-dimensional grid, the tpj blocks of a job next to each other: blocks that run at the same time read neighbouring pieces of
  // the same column runs (DRAM pages) instead of one short run each from many transport blocks
  const DeintJob j = jobs[blockIdx.x / tpj];
  const uint32_t ntiles = (j.rows + DT_ROWS - 1) / DT_ROWS;
  const bool     tiled  = j.cols <= DT_MAXC && j.Qm <= 8 && (j.Qm & 1u) == 0 && ((reinterpret_cast<uintptr_t>(j.q) & 3u) == 0) && ((j.rows * j.Qm) & 1u) == 0;
  const bool     g32ok  = (reinterpret_cast<uintptr_t>(j.g) & 3u) == 0;
  const uint32_t item_w = j.Qm / 2, line_w = j.cols * item_w;  // 32-bit words per item / per matrix row (tiled: Qm even)
  for (uint32_t tix = blockIdx.x % tpj; tix < ntiles; tix += tpj) {
    const uint32_t row0 = tix * DT_ROWS, nr = min((uint32_t)DT_ROWS, j.rows - row0), run = nr * j.Qm;  // LLRs per column of this tile
    bool           fast = false;
    uint32_t       c0   = 0;
    if (tiled) {
      __syncthreads();  // the previous tile has been written out
      // ---- load: a warp per column run, 128-bit loads when the run starts and ends on 16-byte boundaries. No index arithmetic
      //      per element: the first versions divided every word index by run / row lengths and were issue-bound (478 instructions
      //      per thread and tile for 2.25 loads and 2.25 stores; SM instruction throughput 67-73 %, DRAM 25-31 % under ncu).
      const uint32_t tile_sa = (uint32_t)__cvta_generic_to_shared(&tile[0][0]);
      const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31u;
      const bool     vec = (run & 7u) == 0 && ((j.rows * j.Qm) & 7u) == 0 && ((row0 * j.Qm) & 7u) == 0 && (reinterpret_cast<uintptr_t>(j.q) & 15u) == 0;
      for (uint32_t col = wid; col < j.cols; col += 8) {
        const int16_t* src = j.q + ((size_t)col * j.rows + row0) * j.Qm;
        const uint32_t sa  = tile_sa + col * (uint32_t)(PITCH * 2);
        if (vec) {
          for (uint32_t i = lane; i < run / 8; i += 32) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        } else {
          for (uint32_t i = lane; i < run / 2; i += 32) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa + 4u * i), "r"(v) : "memory");
          }
        }
      }
      // number of RI scan indices before the tile, and whether one lies inside it (every thread the same search: broadcast loads)
      const uint32_t s_lo = row0 * j.cols * j.Qm, s_hi = s_lo + nr * j.cols * j.Qm;
      uint32_t       lo = 0, hi = j.nri;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(j.ri_scan + mid) < s_lo) lo = mid + 1; else hi = mid;
      }
      c0   = lo;
      fast = g32ok && (c0 & 1u) == 0 && (c0 >= j.nri || __ldg(j.ri_scan + c0) >= s_hi);
      __syncthreads();
      if (fast) {
        // ---- store: the tile's part of g is nr consecutive lines of line_w words. A thread owns the same word positions of
        //      every line it writes, so their places in the tile are worked out once; lines are walked with a fixed stride.
        uint32_t*      dst   = reinterpret_cast<uint32_t*>(j.g) + (s_lo - c0) / 2;
        const uint32_t rowb  = j.Qm * 2;  // bytes per matrix row within a column run
        const bool     first = s_lo == c0;  // this tile holds g[0]: it is q[p_star] (see above)
        if ((line_w & 3u) == 0 && line_w >= 4 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
          const uint32_t per = line_w / 4, jq = threadIdx.x % per, rr = threadIdx.x / per, rpp = 256 / per;  // 128-bit pieces per line
          uint32_t       off[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t w = 4 * jq + i, col = w / item_w, k = w - col * item_w;
            off[i] = tile_sa + col * (uint32_t)(PITCH * 2) + 4u * k;
          }
          if (rr < rpp) {
            for (uint32_t r = rr; r < nr; r += rpp) {
              uint32_t v[4];
#pragma unroll
              for (int i = 0; i < 4; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[i]) : "r"(off[i] + r * rowb));
              if (first && r == 0 && jq == 0) v[0] = (v[0] & 0xffff0000u) | (uint16_t)j.q[j.p_star];
              reinterpret_cast<uint4*>(dst)[r * per + jq] = make_uint4(v[0], v[1], v[2], v[3]);
            }
          }
        } else {
          // a warp per line, lane l the words l, l + 32, ...
          for (uint32_t w = lane; w < line_w; w += 32) {
            const uint32_t col = w / item_w, k = w - col * item_w, off = tile_sa + col * (uint32_t)(PITCH * 2) + 4u * k;
            for (uint32_t r = wid; r < nr; r += 8) {
              uint32_t v;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(off + r * rowb));
              if (first && r == 0 && w == 0) v = (v & 0xffff0000u) | (uint16_t)j.q[j.p_star];
              dst[r * line_w + w] = v;
            }
          }
        }
        continue;
      }
    }
    for (uint32_t it = threadIdx.x; it < nr * j.cols; it += 256) {
      const uint32_t r = it / j.cols, col = it - r * j.cols, row = row0 + r;
      const uint32_t s0  = (row * j.cols + col) * j.Qm;  // scan index of the item's first LLR
      const int16_t* src = tiled ? &tile[col][r * j.Qm] : j.q + (size_t)row * j.Qm + (size_t)col * j.rows * j.Qm;
      // number of RI scan indices < s0 (binary search), then walk the list together with the Qm values
      uint32_t lo = 0, hi = j.nri;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (j.ri_scan[mid] < s0) lo = mid + 1; else hi = mid;
      }
      const bool clean = (lo >= j.nri || j.ri_scan[lo] >= s0 + j.Qm) && ((s0 - lo) & 1u) == 0 && s0 != lo && tiled && g32ok;
      if (clean) {
        // no RI inside the item, even rank, not the item that owns g[0]: Qm / 2 aligned 32-bit stores
        uint32_t* dst = reinterpret_cast<uint32_t*>(j.g + (s0 - lo));
        for (uint32_t w2 = 0; w2 < j.Qm / 2; w2++) dst[w2] = *reinterpret_cast<const uint32_t*>(src + 2 * w2);
        continue;
      }
      for (uint32_t bit = 0; bit < j.Qm; bit++) {
        const uint32_t s = s0 + bit;
        if (lo < j.nri && j.ri_scan[lo] == s) {
          lo++;
          continue;
        }
        const uint32_t rank = s - lo;
        j.g[rank] = rank == 0 ? j.q[j.p_star] : src[bit];
      }
    }
  }
}

/*
 * Many small host<->device transfers as ONE launch: the job list and the host buffers are page-locked memory that the
 * GPU addresses directly (unified virtual addressing), so block (x, j) simply copies its share of job j. A transport-block
 * submission moves one buffer per TB (e-bits in, bytes out) and, without device-resident soft buffers, two per code
 * block; as cudaMemcpyAsync calls those serialise on the driver (~4 us each, process-wide lock) and bound the
 * multi-threaded uplink case, as one kernel they cost one launch.
 */
struct CopyJob {
  const void* src;
  void*       dst;
  uint64_t    bytes;
};
__global__ void __launch_bounds__(256) gather_copy_kernel(const CopyJob* __restrict__ jobs)
{
  const CopyJob  j   = jobs[blockIdx.y];
  const uint64_t tid = (uint64_t)blockIdx.x * 256 + threadIdx.x, nth = (uint64_t)gridDim.x * 256;
  const uint8_t* s   = static_cast<const uint8_t*>(j.src);
  uint8_t*       d   = static_cast<uint8_t*>(j.dst);
  if ((((uintptr_t)s | (uintptr_t)d) & 15u) == 0) {
    const uint64_t n16 = j.bytes / 16;
    for (uint64_t i = tid; i < n16; i += nth) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
    for (uint64_t i = n16 * 16 + tid; i < j.bytes; i += nth) d[i] = s[i];
  } else if ((((uintptr_t)s | (uintptr_t)d) & 3u) == 0) {
    const uint64_t n4 = j.bytes / 4;
    for (uint64_t i = tid; i < n4; i += nth) reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
    for (uint64_t i = n4 * 4 + tid; i < j.bytes; i += nth) d[i] = s[i];
  } else {
    for (uint64_t i = tid; i < j.bytes; i += nth) d[i] = s[i];
  }
}

}  // namespace srsb200

namespace srsb200 {
/*
 * Soft demodulation to int16 LLRs: srsran_demod_soft_demodulate_s (lib/src/phy/modem/demod_soft.c:871-894) as the reference's
 * AVX2 / SSE build computes it, bit for bit - including WHICH symbols take which rounding: the 16QAM / 64QAM SSE bodies convert
 * four symbols per trip with round-to-nearest-even and a saturating pack (:250-287, :569-629), the last nsymbols % 4 symbols go
 * through scalar code that truncates (the 16QAM tail subtracts its threshold in floating point first, :290-298); QPSK truncates
 * everywhere but only the 16 values of an AVX2 trip are saturated (srsran_vec_convert_fi, vector_simd.c:436-472, simd.h:1866-1871);
 * BPSK is double arithmetic, 256QAM float arithmetic with truncation. One thread per symbol, coalesced 8-byte loads, the Qm LLRs
 * of a symbol are one contiguous store run. HBM-bound: 8 bytes in, 2 Qm bytes out per symbol.
 */
struct DemodJob {
  const float* sym;   // (re, im) pairs
  int16_t*     llr;   // nsym * bits per symbol
  uint32_t     nsym;
  uint32_t     mod;   // srsran_mod_t: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM
};
// _mm_cvtps_epi32 / _mm_cvttps_epi32: out of range and NaN give the "integer indefinite" value 0x80000000
__device__ __forceinline__ int32_t dm_cvt_rne(float v) { return (v > -2147483904.0f && v < 2147483648.0f) ? __float2int_rn(v) : INT32_MIN; }
__device__ __forceinline__ int32_t dm_cvt_trunc(float v) { return (v > -2147483904.0f && v < 2147483648.0f) ? __float2int_rz(v) : INT32_MIN; }
__device__ __forceinline__ int16_t dm_packs(int32_t v) { return (int16_t)min(32767, max(-32768, v)); }
__device__ __forceinline__ int16_t dm_abs16(int16_t v) { return (int16_t)(v < 0 ? (uint16_t)(0u - (uint16_t)v) : (uint16_t)v); }  // |-32768| = -32768
// the C cast (short)(float) of the scalar tails on x86: cvttss2si to 32 bits, then the low half
__device__ __forceinline__ int16_t dm_cast16(float v) { return (int16_t)dm_cvt_trunc(v); }

__global__ void __launch_bounds__(256) demod_kernel(const DemodJob* __restrict__ jobs)
{
  const DemodJob j = jobs[blockIdx.y];
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < j.nsym; i += gridDim.x * 256) {
    const float2 s = *reinterpret_cast<const float2*>(j.sym + 2 * (size_t)i);
    const float  v[2] = {s.x, s.y};
    switch (j.mod) {
      case 0: {
        const double d = (double)__fmul_rn(-100.0f, __fadd_rn(s.x, s.y)) * 0.70710678118654752440;  // -SCALE * (re + im) in float, * M_SQRT1_2 in double
        j.llr[i] = (d > -2147483649.0 && d < 2147483648.0) ? (int16_t)__double2int_rz(d) : (int16_t)0;
      } break;
      case 1: {
        const float    scale = (float)(-100 * 1.41421356237309504880);
        const uint32_t len = 2 * j.nsym, nsimd = len - len % 16;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const float t = __fmul_rn(v[c], scale);
          j.llr[2 * (size_t)i + c] = (2 * i + c < nsimd) ? dm_packs(dm_cvt_trunc(t)) : dm_cast16(t);
        }
      } break;
      case 2: {
        const int16_t offset = 252;  // (short)(2 * 400 / sqrtf(10))
        const bool    simd   = i < 4 * (j.nsym / 4);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          int16_t l0, l1;
          if (simd) {
            l0 = dm_packs(dm_cvt_rne(__fmul_rn(v[c], -400.0f)));
            l1 = (int16_t)(uint16_t)((uint16_t)dm_abs16(l0) - (uint16_t)offset);
          } else {
            const int16_t y = dm_cast16(__fmul_rn(400.0f, v[c]));
            l0 = (int16_t)-y;
            l1 = dm_cast16(__fsub_rn((float)abs((int)y), __fdiv_rn(800.0f, __fsqrt_rn(10.0f))));  // abs(yre) - 2 * 400 / sqrtf(10): int - float
          }
          j.llr[4 * (size_t)i + c]     = l0;
          j.llr[4 * (size_t)i + 2 + c] = l1;
        }
      } break;
      case 3: {
        const int16_t off1 = 432, off2 = 216;  // (short)(4 * 700 / sqrtf(42)), (short)(2 * 700 / sqrtf(42))
        const bool    simd = i < 4 * (j.nsym / 4);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          int16_t y, a1, a2;
          if (simd) {
            y  = dm_packs(dm_cvt_rne(__fmul_rn(v[c], -700.0f)));
            a1 = (int16_t)(uint16_t)((uint16_t)dm_abs16(y) - (uint16_t)off1);
            a2 = (int16_t)(uint16_t)((uint16_t)dm_abs16(a1) - (uint16_t)off2);
          } else {
            const int16_t t = dm_cast16(__fmul_rn(700.0f, v[c]));
            y  = (int16_t)-t;
            a1 = (int16_t)((int16_t)abs((int)t) - off1);
            a2 = (int16_t)((int16_t)abs((int)a1) - off2);
          }
          j.llr[6 * (size_t)i + c]     = y;
          j.llr[6 * (size_t)i + 2 + c] = a1;
          j.llr[6 * (size_t)i + 4 + c] = a2;
        }
      } break;
      default: {
        // 256QAM: real = fabsf(real) - 8 / sqrtf(170) ... in float, every LLR = (short)(1000 * real)
        const float sq = __fsqrt_rn(170.0f), c8 = __fdiv_rn(8.0f, sq), c4 = __fdiv_rn(4.0f, sq), c2 = __fdiv_rn(2.0f, sq);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float r = -v[c];
          j.llr[8 * (size_t)i + c]     = dm_cast16(__fmul_rn(1000.0f, r));
          r = __fsub_rn(fabsf(r), c8);
          j.llr[8 * (size_t)i + 2 + c] = dm_cast16(__fmul_rn(1000.0f, r));
          r = __fsub_rn(fabsf(r), c4);
          j.llr[8 * (size_t)i + 4 + c] = dm_cast16(__fmul_rn(1000.0f, r));
          r = __fsub_rn(fabsf(r), c2);
          j.llr[8 * (size_t)i + 6 + c] = dm_cast16(__fmul_rn(1000.0f, r));
        }
      } break;
    }
  }
}
}  // namespace srsb200

namespace srsb200 {
/*
 * Descrambling of int16 LLRs in place (srsran_sequence_apply_s, lib/src/phy/common/sequence.c:507-561, as called by
 * srsran_sequence_pusch_apply_s / _pdsch_apply_s, lib/src/phy/phch/sequences.c:95-146): q[i] = c(i) ? -q[i] : q[i] with int16 wrap.
 * A thread owns 32 consecutive positions: it jumps the two LFSR windows of the Gold sequence to its word (GF(2) matrix powers,
 * like rm_rx_kernel<true>) and steps them 32 times. x1 / x2 = windows at sequence position 0 (after Nc = 1600).
 */
struct ScrJob {
  int16_t* q;
  uint32_t n, x1, x2;
};
__global__ void __launch_bounds__(256) descramble_kernel(const ScrJob* __restrict__ jobs, const uint32_t* __restrict__ gold)
{
  const ScrJob   j    = jobs[blockIdx.y];
  const uint32_t word = blockIdx.x * 256 + threadIdx.x;
  if (32ull * word >= j.n) return;
  uint32_t w1 = gold_jump(gold, j.x1, 32 * word), w2 = gold_jump(gold + 32 * GOLD_POWERS, j.x2, 32 * word);
  const uint32_t end = min(j.n, 32 * word + 32);
  for (uint32_t i = 32 * word; i < end; i++) {
    if ((w1 ^ w2) & 1u) j.q[i] = (int16_t)(uint16_t)(0u - (uint16_t)j.q[i]);
    w1 = (w1 >> 1) | ((uint32_t)(__popc(w1 & 0x9u) & 1) << 30);
    w2 = (w2 >> 1) | ((uint32_t)(__popc(w2 & 0xFu) & 1) << 30);
  }
}
// out[i] = q[pos[i]]: the handful of (descrambled) LLRs the host-side UCI decoding reads (RI / ACK positions)
struct GatherJob {
  const int16_t*  q;
  const uint32_t* pos;
  int16_t*        out;
  uint32_t        n, qlen;
};
__global__ void __launch_bounds__(256) gather16_kernel(const GatherJob* __restrict__ jobs)
{
  const GatherJob j = jobs[blockIdx.y];
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < j.n; i += gridDim.x * 256) {
    const uint32_t p = j.pos[i];
    j.out[i] = p < j.qlen ? j.q[p] : (int16_t)0;
  }
}
}  // namespace srsb200
