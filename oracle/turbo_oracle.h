/*
 * turbo_oracle.h - TEST INFRASTRUCTURE ONLY (see turbo_oracle.c).
 *
 * Clean-room CPU restatement of the srsRAN 4G LTE turbo-decode hot path (generic int16 decoder, natural layout):
 * the bit-exact checker for the CUDA engine. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library. Parity status: PINNED - checked against the compiled reference (oracle/_ref) over all 188 block
 * sizes and against the reference's own known-answer vectors (tests/test_oracle_*.py, tests/golden/).
 */
#ifndef TURBO_ORACLE_H
#define TURBO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NOF_CB_SIZES 188
#define ORC_MAX_K 6144
#define ORC_SOFTBUFFER_SIZE 18600
#define ORC_CRC24A 0x1864CFBu
#define ORC_CRC24B 0x1800063u

/* code-block size table and QPP interleaver */
int orc_cbsize(uint32_t idx);            /* K or -1 */
int orc_cbindex(uint32_t K);             /* first idx with size >= K, or -1 */
int orc_cbindex_exact(uint32_t K);       /* idx with size == K, or -1 */
int orc_qpp(uint32_t K, uint16_t* fwd, uint16_t* rev);
int orc_cbsegm(uint32_t tbs, uint32_t* out12); /* F C K1 K2 K1_idx K2_idx C1 C2 tbs L_tb L_cb Z ; 0 ok / -1 */

/* CRC, MSB first, init 0 */
uint32_t orc_crc_bytes(uint32_t poly, int order, const uint8_t* data, int nbits);
uint32_t orc_crc_bits(uint32_t poly, int order, const uint8_t* bits, int nbits);

/* encoder side (test-vector generation) */
int orc_tcod_encode(const uint8_t* in_bits, uint8_t* out_bits, uint32_t K);
int orc_rm_tx(const uint8_t* coded_bits, uint32_t K, uint8_t* e_bits, uint32_t E, uint32_t rv);

/* rate de-matching */
int orc_rm_table(uint32_t cb_idx, uint32_t rv, uint16_t* table);
int orc_rm_rx(const int16_t* e, int16_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv);

/* decoder */
int orc_map_gen(uint32_t K, const int16_t* input, const int16_t* app, const int16_t* parity, int16_t* output);
int orc_tdec_trace(uint32_t K, const int16_t* in, uint32_t nof_iter, uint8_t* out_bytes, int16_t* dump);
int orc_tdec_run_all(uint32_t K, const int16_t* in, uint32_t nof_iter, uint8_t* out_bytes);
double orc_tdec_batch(uint32_t K, const int16_t* in, uint32_t n, uint32_t max_iter, int early_stop, int nthreads,
                      uint8_t* out, uint8_t* noi, uint8_t* crc_ok);

/* transport-block loop */
int orc_decode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const int16_t* e_bits,
                  uint32_t max_iterations, int16_t* buffer_f, uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc,
                  uint8_t* data, uint32_t* cb_noi, float* avg_iterations);

/* transport-block encode (bit-level restatement of encode_tb_off, sch.c:240-358). data: tbs/8 packed bytes;
 * e_bits: packed MSB-first, (nof_e_bits+7)/8 bytes, zero-filled then written up to Qm*(nof_e_bits/Qm) bits.
 * Returns 0, -1 (filler bits / Qm == 0 / bad tbs) or -2 (null pointers). */
int orc_encode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const uint8_t* data, uint8_t* e_bits);

/* UL-SCH channel de-interleaver (36.212 5.2.2.8) as realised by ulsch_deinterleave + ulsch_interleave_gen +
 * srsran_vec_lut_sis (sch.c:661-682,994-1021, vector.c:147-152): g[lut[p]] = q[p] for p ascending, lut[p] = running
 * count of non-RI positions in row-major (row j, column i, bit k) order, position p = j*Qm + i*rows*Qm + k, and 0 for
 * positions that carry RI (so g[0] ends up holding the LAST position that maps to 0). g must hold H_prime_total*Qm. */
int orc_ulsch_deinterleave(const int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits,
                           const uint32_t* ri_positions, uint32_t nof_ri_bits);

/* LTE scrambling sequence (36.211 7.2: Gold sequence of two length-31 LFSRs, Nc = 1600) applied as a sign to int16 LLRs:
 * out[i] = c(i) ? -in[i] : in[i] (int16 wrap), what srsran_sequence_apply_s does (lib/src/phy/common/sequence.c:507-561);
 * orc_sequence_bits writes c(0..len-1) as 0/1 bytes. */
void orc_sequence_bits(uint32_t c_init, uint8_t* c, uint32_t len);
void orc_sequence_apply_s(const int16_t* in, int16_t* out, uint32_t len, uint32_t c_init);

/* soft demodulation to int16 LLRs as srsran_demod_soft_demodulate_s computes it in the AVX2 build (SIMD body rounds to nearest even
 * and saturates, the scalar tails truncate); mod 0..4 = BPSK, QPSK, 16QAM, 64QAM, 256QAM; symbols = (re, im) float pairs */
int orc_demod_soft_demodulate_s(int mod, const float* symbols, int16_t* llr, int nsymbols);

/* ------------------------------------------------------------------ 8-bit LLR mode (turbo_oracle8.c)
 * The reference's windowed saturating int8 decoders (turbodecoder_win.h with llr_t = int8_t) restated in natural order.
 * orc_tdec8_windows: 32 / 16 = number of windows the reference's AUTO mode uses for this K, 0 = not decoded in 8-bit
 * arithmetic there (K <= 800 or K not a multiple of 16). */
uint32_t orc_tdec8_windows(uint32_t K);
int orc_tdec8_trace(uint32_t K, const int8_t* in, uint32_t nof_iter, uint8_t* out_bytes, int8_t* dump);
double orc_tdec8_batch(uint32_t K, const int8_t* in, uint32_t n, uint32_t max_iter, int early_stop, int nthreads, uint8_t* out, uint8_t* noi,
                       uint8_t* crc_ok);
int orc_rm_rx8(const int8_t* e, int8_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv);
int orc_decode_tb8(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const int8_t* e_bits, uint32_t max_iterations, int8_t* buffer_b,
                   uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc, uint8_t* data, uint32_t* cb_noi, float* avg_iterations);

#ifdef __cplusplus
}
#endif
#endif
