/*
 * rm_kernels.cuh - rate de-matching + HARQ soft combining (srsran_rm_turbo_rx_lut, lib/src/phy/fec/turbo/rm_turbo.c:390-445)
 * and transport-block CRC24A (decode_tb, lib/src/phy/phch/sch.c:563) as sm_100a kernels.
 *
 * The reference computes  output[T[i mod L]] += input[i]  for i < E  (L = 3K+12, int16 wrap). T is a permutation of
 * 0..L-1, so thread n < L owns every received LLR that lands on soft-buffer position T[n]:
 *     sum_n = e[n] + e[n+L] + e[n+2L] + ...   (repetition when E > L),   buf[T[n]] += sum_n
 * Loads of e are coalesced, the read-modify-write of the soft buffer is conflict-free without atomics, and the result is
 * the reference's bit for bit because 16-bit wrapping addition is associative.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

struct RmJob {
  const int16_t*  e;      // first received LLR of this code block (device)
  int16_t*        buf;    // soft buffer of this code block, natural layout (device)
  const uint16_t* table;  // T[0..L)
  uint32_t        E;
  uint32_t        L;
};

// grid = (ceil(maxL/256), n_jobs), block = 256
__global__ void __launch_bounds__(256) rm_rx_kernel(const RmJob* __restrict__ jobs)
{
  const RmJob    j = jobs[blockIdx.y];
  const uint32_t n = blockIdx.x * 256 + threadIdx.x;
  if (n >= j.L || n >= j.E) return;
  uint32_t sum = 0;
  for (uint32_t i = n; i < j.E; i += j.L) sum += (uint16_t)j.e[i];
  const uint32_t t = j.table[n];
  j.buf[t] = (int16_t)(uint16_t)((uint16_t)j.buf[t] + sum);
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
  for (uint32_t byte = blockIdx.x * 256 + threadIdx.x; byte < j.nbits / 8; byte += gridDim.x * 256) {
    uint32_t v = j.data[byte];
    while (v) {
      int      b   = 31 - __clz(v);        // bit b of the byte (b = 7 is the earliest bit)
      uint32_t pos = byte * 8 + (7 - b);   // position in the message
      acc ^= words[j.nbits - 1 - pos];
      v &= ~(1u << b);
    }
  }
  acc = __reduce_xor_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(&crc_out[blockIdx.y], acc);
}

/*
 * UL-SCH channel de-interleaver (36.212 5.2.2.8; ulsch_deinterleave, lib/src/phy/phch/sch.c:994-1021). The reference
 * builds lut[] with a running counter over the rows x cols x Qm matrix in (row, column, bit) order - position
 * p = row*Qm + col*rows*Qm + bit - skipping the positions that carry RI (lut = 0 there), then scatters g[lut[p]] = q[p]
 * for ascending p. Here thread s owns the s-th matrix element in that order: its rank is s minus the number of RI elements
 * before it (binary search in the job's sorted list), so writes are coalesced and no table is built. g[0] receives
 * q[p_star], the last position the sequential loop would have written there (an RI position when there is one).
 */
struct DeintJob {
  const int16_t*  q;
  int16_t*        g;
  const uint32_t* ri_scan;  // sorted scan-order indices of the RI positions
  uint32_t        nri, rows, cols, Qm, p_star;
};
__global__ void __launch_bounds__(256) ulsch_deint_kernel(const DeintJob* __restrict__ jobs)
{
  const DeintJob j = jobs[blockIdx.y];
  const uint32_t n = j.rows * j.cols * j.Qm;
  for (uint32_t s = blockIdx.x * 256 + threadIdx.x; s < n; s += gridDim.x * 256) {
    // number of RI scan indices < s, and whether s itself is one
    uint32_t lo = 0, hi = j.nri;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (j.ri_scan[mid] < s) lo = mid + 1; else hi = mid;
    }
    if (lo < j.nri && j.ri_scan[lo] == s) continue;
    const uint32_t rank = s - lo;
    const uint32_t row = s / (j.cols * j.Qm), rem = s - row * j.cols * j.Qm, col = rem / j.Qm, bit = rem - col * j.Qm;
    const uint32_t p = row * j.Qm + col * j.rows * j.Qm + bit;
    j.g[rank] = j.q[rank == 0 ? j.p_star : p];
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
