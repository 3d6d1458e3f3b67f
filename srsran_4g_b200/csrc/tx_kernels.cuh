/*
 * tx_kernels.cuh - transport-block ENCODE direction as one sm_100a kernel per code block:
 *   CRC attach (TB CRC24A on the last block, CRC24B per block when C > 1)    srsran_tcod_encode_lut, turbocoder.c:186-258
 *   turbo encode: two 8-state RSC encoders, QPP interleaver, trellis tails    turbocoder.c:260-360
 *   rate matching: sub-block interleave + circular buffer + bit selection     srsran_rm_turbo_tx_lut, rm_turbo.c:345-388
 *   bit-packed output at the block's bit offset in the transport block        encode_tb_off, sch.c:240-358
 *
 * The reference runs each encoder as a byte-wise state machine, serial over the K bits. Here one block of 384 threads owns a
 * code block and a thread owns one 32-bit word of one encoder's input. The RSC recursion is linear over GF(2), so
 *   pass 1: every thread runs its word from state 0 and records the end state e_w
 *   scan  : start state s_{w+1} = A^32 s_w ^ e_w  (A = zero-input state transition; feedback 1+D^2+D^3 is primitive, A^7 = I)
 *   pass 2: every thread re-runs its word from the true start state and emits the parity word
 * Rate matching is the same permutation table T the receive side uses (e[n] = d[T[n mod L]]), gathered from the three
 * bit streams held in shared memory and OR-ed into the zeroed output at bit granularity.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srsb200 {

static constexpr int TX_THREADS = 384;           // 2 encoders x 192 words (K <= 6144)
static constexpr int TX_MAXW    = 192;

struct TxJob {
  const uint8_t*  src;     // payload bytes of this code block inside the device copy of the transport block
  const uint16_t* table;   // T[0..3K+12) of (K, rv)
  uint32_t*       out;     // transport block's packed e-bits (device, 4-byte aligned, zeroed)
  uint32_t        n_src;   // payload bytes taken from src
  uint32_t        tb;      // index into tb_crc[]
  uint32_t        flags;   // bit 0: last block (TB CRC appended), bit 1: CRC24B appended
  uint32_t        K, f1, f2;
  uint32_t        E;       // bits to transmit
  uint32_t        wp;      // bit offset of this block in the transport block's e-bits
};

// one step of the constituent encoder (36.212 5.1.3.2.1: g0 = 1+D^2+D^3 feedback, g1 = 1+D+D^3); s = d0 | d1<<1 | d2<<2
__device__ __forceinline__ uint32_t rsc_step(uint32_t& s, uint32_t u)
{
  const uint32_t d0 = s & 1, d1 = (s >> 1) & 1, d2 = (s >> 2) & 1;
  const uint32_t a = u ^ d1 ^ d2;
  s = a | (d0 << 1) | (d1 << 2);
  return a ^ d0 ^ d2;
}
// termination step (switch in the lower position): input = feedback, so a = 0; returns x | z<<1
__device__ __forceinline__ uint32_t rsc_tail(uint32_t& s)
{
  const uint32_t d0 = s & 1, d1 = (s >> 1) & 1, d2 = (s >> 2) & 1;
  const uint32_t x = d1 ^ d2, z = d0 ^ d2;
  s = (d0 << 1) | (d1 << 2);
  return x | (z << 1);
}

__device__ __forceinline__ uint32_t stream_bit(const uint32_t* w, uint32_t i) { return (w[i >> 5] >> (31 - (i & 31))) & 1u; }

__global__ void __launch_bounds__(TX_THREADS) tx_cb_kernel(const TxJob* __restrict__ jobs, const uint32_t* __restrict__ tb_crc,
                                                           const uint32_t* __restrict__ crc24b_words)
{
  __shared__ uint32_t sw[3][TX_MAXW + 1];   // systematic / parity 1 / parity 2, bit i at word i>>5, position 31-(i&31)
  __shared__ uint32_t xin[TX_MAXW];         // interleaved input of encoder 2
  __shared__ uint8_t  est[2][TX_MAXW];      // pass 1 end states, then true start states
  __shared__ uint8_t  cbb[4 * TX_MAXW + 4];
  __shared__ uint32_t tail[12];
  __shared__ uint32_t crc_acc;

  const TxJob    j   = jobs[blockIdx.x];
  const uint32_t tid = threadIdx.x, K = j.K, nw = (K + 31) / 32;

  // ---- 1. assemble the block's bytes: payload [+ TB CRC] [+ CB CRC]
  if (tid == 0) crc_acc = 0;
  for (uint32_t b = tid; b < 4 * nw + 4; b += TX_THREADS) cbb[b] = b < j.n_src ? j.src[b] : 0;
  __syncthreads();
  uint32_t have = j.n_src;
  if (j.flags & 1u) {
    if (tid < 3) cbb[have + tid] = (uint8_t)(tb_crc[j.tb] >> (16 - 8 * tid));
    have += 3;
    __syncthreads();
  }
  if (j.flags & 2u) {
    // CRC24B by linearity: XOR of x^(m+24) mod g over the set bits, m = distance from the end of the message
    const uint32_t nbits = have * 8;
    uint32_t acc = 0;
    for (uint32_t b = tid; b < have; b += TX_THREADS) {
      uint32_t v = cbb[b];
      while (v) {
        const int bit = 31 - __clz(v);
        acc ^= crc24b_words[nbits - 1 - (b * 8 + (7 - bit))];
        v &= ~(1u << bit);
      }
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    if ((tid & 31) == 0 && acc) atomicXor(&crc_acc, acc);
    __syncthreads();
    if (tid < 3) cbb[have + tid] = (uint8_t)(crc_acc >> (16 - 8 * tid));
    __syncthreads();
  }
  // have + 3*(CRC24B) == nb by construction of the job

  // ---- 2. systematic words, interleaved words
  if (tid < nw) sw[0][tid] = ((uint32_t)cbb[4 * tid] << 24) | ((uint32_t)cbb[4 * tid + 1] << 16) | ((uint32_t)cbb[4 * tid + 2] << 8) | cbb[4 * tid + 3];
  __syncthreads();
  if (tid < nw) {
    const uint32_t i0 = 32 * tid;
    uint32_t p = (uint32_t)(((uint64_t)j.f1 * i0 + (uint64_t)j.f2 * i0 % K * i0) % K);      // pi(i0)
    uint32_t g = (uint32_t)((j.f1 + j.f2 + 2ull * j.f2 * i0) % K);                           // pi(i+1) - pi(i)
    const uint32_t g2 = (2 * j.f2) % K;
    uint32_t word = 0;
    const uint32_t n = min(32u, K - i0);
    for (uint32_t b = 0; b < n; b++) {
      word |= stream_bit(sw[0], p) << (31 - b);
      p += g;  if (p >= K) p -= K;
      g += g2; if (g >= K) g -= K;
    }
    xin[tid] = word;
  }
  __syncthreads();

  // ---- 3. encoders: threads [0,192) encoder 1, [192,384) encoder 2
  const uint32_t enc = tid >= TX_MAXW, w = tid - enc * TX_MAXW;
  const uint32_t nbw = w < nw ? min(32u, K - 32 * w) : 0;
  const uint32_t inw = w < nw ? (enc ? xin[w] : sw[0][w]) : 0;
  if (w < nw) {
    uint32_t s = 0;
    for (uint32_t b = 0; b < nbw; b++) rsc_step(s, (inw >> (31 - b)) & 1u);
    est[enc][w] = (uint8_t)s;
  }
  __syncthreads();
  if (w == 0) {
    // A^32 = A^4 as a packed 8-entry table (3 bits per entry)
    uint32_t lut = 0;
    for (uint32_t s0 = 0; s0 < 8; s0++) {
      uint32_t s = s0;
      for (int k = 0; k < 32 % 7; k++) {
        const uint32_t d0 = s & 1, d1 = (s >> 1) & 1, d2 = (s >> 2) & 1;
        s = (d1 ^ d2) | (d0 << 1) | (d1 << 2);   // zero input: a = fb
      }
      lut |= s << (3 * s0);
    }
    uint32_t s = 0;
    for (uint32_t k = 0; k < nw; k++) {
      const uint32_t e = est[enc][k];
      est[enc][k] = (uint8_t)s;
      if (k + 1 < nw) s = ((lut >> (3 * s)) & 7u) ^ e;
    }
  }
  __syncthreads();
  if (w < nw) {
    uint32_t s = est[enc][w], z = 0;
    for (uint32_t b = 0; b < nbw; b++) z |= rsc_step(s, (inw >> (31 - b)) & 1u) << (31 - b);
    sw[1 + enc][w] = z;
    if (w == nw - 1) {
      // trellis termination: 3 steps, x and z of each (turbocoder.c:318-352)
      for (int k = 0; k < 3; k++) {
        const uint32_t xz = rsc_tail(s);
        tail[6 * enc + 2 * k]     = xz & 1u;
        tail[6 * enc + 2 * k + 1] = xz >> 1;
      }
    }
  }
  __syncthreads();

  // ---- 4. rate matching + packing: thread -> one byte of this block's e-bits at a time
  const uint32_t L = 3 * K + 12, sh = j.wp & 7u, B0 = j.wp >> 3;
  for (uint32_t b = tid; b < (j.E + 7) / 8; b += TX_THREADS) {
    uint32_t m = (8 * b) % L, v = 0;
    const uint32_t n = min(8u, j.E - 8 * b);
    for (uint32_t q = 0; q < n; q++) {
      const uint32_t idx = j.table[m];
      uint32_t bit;
      if (idx < 3 * K) {
        const uint32_t k = idx / 3, s = idx - 3 * k;
        bit = stream_bit(sw[s], k);
      } else {
        bit = tail[idx - 3 * K];
      }
      v |= bit << (7 - q);
      if (++m == L) m = 0;
    }
    // byte b of the block starts at bit wp + 8b of the transport block: split over two output bytes when wp % 8 != 0
    const uint32_t B  = B0 + b;
    const uint32_t hi = v >> sh, lo = (v << (8 - sh)) & 0xffu;
    if (hi) atomicOr(&j.out[B >> 2], hi << (8 * (B & 3)));
    if (sh && lo) atomicOr(&j.out[(B + 1) >> 2], lo << (8 * ((B + 1) & 3)));
  }
}

}  // namespace srsb200
