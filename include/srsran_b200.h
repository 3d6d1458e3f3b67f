/*
 * srsran_b200.h - C ABI of the B200-native LTE turbo-decode engine (libsrsran_b200.so).
 *
 * Plain pointers and sizes only (no torch / C++ types). Two layers:
 *
 *  (1) Engine + batched entry points (NEW - the reference has no batched equivalent; its decode_tb_cb is a serial
 *      per-code-block loop, lib/src/phy/phch/sch.c:390): every code block of a transport block, or of many transport
 *      blocks across subframes and UEs, is submitted as ONE device launch.
 *
 *  (2) Per-object entry points with the reference's argument meaning and error behaviour, one per reference function
 *      on the hot path. The reference-side binding (integration/srsran_b200_shim.c, built inside the reference tree
 *      with the reference's own struct definitions) maps srsran_tdec_* / srsran_rm_turbo_rx_lut / decode_tb /
 *      srsran_dlsch_decode2 / srsran_ulsch_decode onto them; see INTEGRATION.md.
 *
 * Results (hard bits, per-block half-iteration counts, CRC verdicts, soft-buffer contents) are bit-exact against the
 * reference's generic int16 decoder (lib/src/phy/fec/turbo/turbodecoder_gen.c) with natural-order input.
 *
 * Error codes follow lib/include/srsran/config.h:57-59: 0 success, -1 error, -2 invalid inputs. There is NO CPU
 * fallback: every compute entry point fails with SRSB200_ERROR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef SRSRAN_B200_H
#define SRSRAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSB200_SUCCESS 0
#define SRSB200_ERROR -1
#define SRSB200_ERROR_INVALID_INPUTS -2
#define SRSB200_ERROR_NO_DEVICE -3

#define SRSB200_NOF_CB_SIZES 188     /* lib/include/srsran/phy/fec/cbsegm.h: SRSRAN_NOF_TC_CB_SIZES */
#define SRSB200_MAX_K 6144           /* SRSRAN_TCOD_MAX_LEN_CB, turbodecoder.h:44 */
#define SRSB200_SOFTBUFFER_SIZE 18600 /* SOFTBUFFER_SIZE, lib/include/srsran/phy/fec/softbuffer.h:56 */
#define SRSB200_MIN_TDEC_ITERS 2     /* SRSRAN_PDSCH_MIN_TDEC_ITERS, sch.c:35 */
#define SRSB200_MAX_TDEC_ITERS 10    /* SRSRAN_PDSCH_MAX_TDEC_ITERS, sch.c:36 */

/* which CRC ends a code block: multi-CB transport blocks use CRC24B per CB, single-CB ones CRC24A (sch.c:438-444) */
#define SRSB200_CRC_NONE 0
#define SRSB200_CRC_24A 1
#define SRSB200_CRC_24B 2

typedef struct srsb200_engine srsb200_engine_t;

/* ------------------------------------------------------------------ engine */
/* device < 0: use the current CUDA device. The engine owns one stream, its device tables and its workspace. */
int  srsb200_engine_create(srsb200_engine_t** e, int device);
void srsb200_engine_destroy(srsb200_engine_t* e);
const char* srsb200_last_error(void);
/* number of kernels launched by this engine since creation (bench.py's gpu_launches claim) */
uint64_t srsb200_engine_launch_count(const srsb200_engine_t* e);
/* per-kernel CUDA-event timing on the launching stream (bench.py's roofline leg). profile_read synchronises and
 * returns, per kernel kind (0 extract, 2 emit, 3 rate de-match, 4 TB CRC, 5 alpha/beta scan, 6 window jobs, 7 transport-block encode), the
 * summed milliseconds and launch counts since the previous read. While profiling is enabled a decode runs as one chain
 * of launches (the sub-batch overlap of srsb200_engine_set_subbatches is off) so the durations are not inflated. */
int srsb200_engine_profile(srsb200_engine_t* e, int enable);
int srsb200_engine_profile_read(srsb200_engine_t* e, double ms[8], uint64_t cnt[8]);
/* the same with room for the kinds added later: 8 UL-SCH de-interleaver, 9 soft demodulation / descrambling / gather, 10 regrouping
 * of unfinished code blocks (with n_kinds = 8 these are reported under 3, 3 and 0 as before) */
int srsb200_engine_profile_read_kinds(srsb200_engine_t* e, double* ms, uint64_t* cnt, uint32_t n_kinds);
/* host-pointer submissions of a contiguous equal-size batch are cut into this many ranges whose H2D / decode / D2H
 * overlap on separate streams (default 8, env SRSB200_SUBBATCHES; 1 = off) */
int srsb200_engine_set_subbatches(srsb200_engine_t* e, int n);
/* test hook for the error paths: the nth scratch-buffer request from now on fails (0 = off; env SRSB200_FAIL_ALLOC at
 * engine creation). A failed submission returns SRSB200_ERROR with every queued transport block's ret = SRSB200_ERROR,
 * nothing in flight, and the engine stays usable. */
int srsb200_engine_inject_alloc_failure(srsb200_engine_t* e, int nth);
/* stream the engine launches on (cudaStream_t), so callers can time with events on the same stream */
void* srsb200_engine_stream(const srsb200_engine_t* e);

/* ------------------------------------------------------------------ host-side metadata (pure integer, no device) */
/* replaces srsran_cbsegm_cbsize / srsran_cbsegm_cbindex (lib/src/phy/fec/cbsegm.c:119-140) */
int srsb200_cbsize(uint32_t index);
int srsb200_cbindex(uint32_t long_cb);
/* replaces srsran_cbsegm (cbsegm.c:62-117). out[8] = F C K1 K2 K1_idx K2_idx C1 C2 */
int srsb200_cbsegm(uint32_t tbs, uint32_t out[8]);
/* replaces srsran_tdec_autoimp_get_subblocks[_8bit] (turbodecoder.c:381-393,410-424): always 0 = natural layout,
 * which makes srsran_rm_turbo_rx_lut and the decoder agree on the generic decoder's input order. */
uint32_t srsb200_tdec_autoimp_get_subblocks(uint32_t long_cb);

/* ------------------------------------------------------------------ batched turbo decode (configs 1, 2, 4) */
/*
 * Decode n code blocks in one submission. Code block i has size K[i] (any of the 188 LTE sizes, mixed freely), its
 * 3*K[i]+12 int16 LLRs in the reference's natural order (turbodecoder_gen.c:246-257) start at llr[llr_offset[i]]
 * (offsets in int16 elements). For each block the reference loop of sch.c:426-456 is run:
 *    do { half-iteration; noi++; } while (noi < max_iter && !(early_stop && noi >= min_iter && crc == 0))
 * crc_kind[i] selects the CRC (NONE never stops early and reports crc_ok = 0). With early_stop == 0 this is
 * srsran_tdec_run_all(h, in, out, max_iter, K) (turbodecoder.c:536-549; max_iter == 0 still runs one half-iteration).
 * Outputs: out_bytes + out_offset[i] receives K[i]/8 bytes, MSB first (tdec_gen_decision_byte); noi[i] the number of
 * half-iterations run; crc_ok[i] = 1 when the CRC over the K[i] decided bits is zero after the last one.
 * Host-pointer variant: copies in, launches, copies out, returns when the results are in the host buffers.
 */
int srsb200_tdec_batch(srsb200_engine_t* e, uint32_t n, const uint32_t* K, const uint8_t* crc_kind, const int16_t* llr,
                       const uint64_t* llr_offset, uint64_t llr_len, uint32_t max_iter, uint32_t min_iter, int early_stop,
                       uint8_t* out_bytes, const uint64_t* out_offset, uint64_t out_len, uint8_t* noi, uint8_t* crc_ok);

/*
 * Device-resident variant for equal-size blocks: d_llr [n][3K+12] int16, d_out [n][K/8], d_noi [n], d_crc_ok [n] are
 * DEVICE pointers; work is enqueued on the engine stream and NOT synchronised (time it with events on
 * srsb200_engine_stream()). plan must come from srsb200_tdec_plan_uniform and can be reused for every launch of the
 * same shape; it owns the device workspace.
 */
typedef struct srsb200_plan srsb200_plan_t;
int  srsb200_tdec_plan_uniform(srsb200_engine_t* e, uint32_t n, uint32_t K, int crc_kind, srsb200_plan_t** plan);
void srsb200_plan_destroy(srsb200_plan_t* plan);
/* Diagnostics: plans of one block size with 64 or more groups regroup their unfinished code blocks during a device-resident
 * decode. Writes, per range of the plan's LAST decode (up to n entries, the rest 0), the half-iteration
 * count after which the range was regrouped, 0 if it was not; synchronises the engine. Returns the number of entries a plan
 * of this kind has (0: this plan never regroups). plan == NULL: the engine's latest transport-block submission (16-bit LLRs). */
int  srsb200_plan_regroup_points(srsb200_engine_t* e, srsb200_plan_t* plan, uint32_t* points, uint32_t n);
int  srsb200_tdec_run_plan_dev(srsb200_engine_t* e, srsb200_plan_t* plan, const int16_t* d_llr, uint32_t max_iter,
                               uint32_t min_iter, int early_stop, uint8_t* d_out, uint8_t* d_noi, uint8_t* d_crc_ok);
/* Device-resident submissions are asynchronous AND pipelined: every plan is bound to one of two lanes of streams, a
 * submission waits for the work already queued on the engine stream and for the previous submission of the SAME plan,
 * and submissions of different plans overlap on the GPU (the latency-bound last half-iterations of one batch hide under
 * the bandwidth-bound first ones of the next). Results are complete after srsb200_engine_sync(); to order other work on
 * srsb200_engine_stream() after them without blocking the host call srsb200_engine_flush() first. Submissions of two
 * plans in flight at once must not share output buffers. */
/* Page-locked host memory for callers that do not link the CUDA runtime themselves: buffers obtained here (or registered
 * with srsb200_host_register) go straight to the copy engines; pageable ones are staged through an engine-owned pinned
 * arena with one extra host memcpy. */
void* srsb200_host_alloc(size_t bytes);
void  srsb200_host_free(void* p);
int   srsb200_host_register(void* p, size_t bytes);   /* cudaHostRegister on an existing allocation */
int   srsb200_host_unregister(void* p);
int  srsb200_engine_flush(srsb200_engine_t* e);
int  srsb200_engine_sync(srsb200_engine_t* e);

/* ------------------------------------------------------------------ per-object decoder: srsran_tdec_* */
/*
 * Handle-based mirror of srsran_tdec_t (lib/include/srsran/phy/fec/turbo/turbodecoder.h:63-116). The device keeps the
 * decoder state (app1/app2/ext1/ext2 equivalents) between srsb200_tdec_iteration calls exactly as the reference keeps
 * them in the object.
 */
typedef struct srsb200_tdec srsb200_tdec_t;
int  srsb200_tdec_init(srsb200_tdec_t** h, srsb200_engine_t* e, uint32_t max_long_cb); /* srsran_tdec_init[_manual] :97-99 */
void srsb200_tdec_free(srsb200_tdec_t* h);                                            /* srsran_tdec_free :101 */
int  srsb200_tdec_new_cb(srsb200_tdec_t* h, uint32_t long_cb);        /* :105  -1 if K > max or not an LTE size */
int  srsb200_tdec_get_nof_iterations(srsb200_tdec_t* h);              /* :107 */
int  srsb200_tdec_iteration(srsb200_tdec_t* h, const int16_t* input, uint8_t* output); /* :113 one half-iteration + decision */
int  srsb200_tdec_run_all(srsb200_tdec_t* h, const int16_t* input, uint8_t* output, uint32_t nof_iterations,
                          uint32_t long_cb);                          /* :115-116 */
/* the hard decision of the latest half-iteration again (the static tdec_decision_byte wrapper, turbodecoder.c:370-378;
 * the north-star's "get_hard_decision"); -1 before the first half-iteration */
int  srsb200_tdec_get_hard_decision(srsb200_tdec_t* h, uint8_t* output);

/* ------------------------------------------------------------------ 8-bit LLR mode: srsran_tdec_*_8bit, srsran_rm_turbo_rx_lut_8bit */
/*
 * The reference's 8-bit mode is not the generic algorithm in fewer bits: srsran_tdec_iteration_8bit (turbodecoder.c:458-484,
 * 551-555) runs the windowed SIMD decoders of turbodecoder_win.h with llr_t = int8_t (saturating arithmetic, 32 or 16
 * windows with 40-step warm-up, max-normalisation, extrinsic >> 1). The engine reproduces them bit for bit (win8_kernels.cuh;
 * the CPU checker of the test suite is pinned to the compiled reference). srsb200_tdec8_windows: number of windows for this K, 0 = the
 * reference does not decode this size in 8-bit arithmetic (it widens to int16 SSE decoders, turbodecoder.c:443-476).
 */
uint32_t srsb200_tdec8_windows(uint32_t K);
/* srsb200_tdec_batch with int8 LLRs (natural order, 3K+12 per block; llr_offset / llr_len in elements = bytes); every K must have
 * srsb200_tdec8_windows(K) != 0 */
int srsb200_tdec_batch8(srsb200_engine_t* e, uint32_t n, const uint32_t* K, const uint8_t* crc_kind, const int8_t* llr, const uint64_t* llr_offset,
                        uint64_t llr_len, uint32_t max_iter, uint32_t min_iter, int early_stop, uint8_t* out_bytes, const uint64_t* out_offset,
                        uint64_t out_len, uint8_t* noi, uint8_t* crc_ok);
/* replaces srsran_rm_turbo_rx_lut_8bit (rm_turbo.c:447-483): output[T[i mod (3K+12)]] += input[i] in wrapping int8, natural layout */
int srsb200_rm_turbo_rx_lut8(srsb200_engine_t* e, const int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);

/* ------------------------------------------------------------------ rate de-matching: srsran_rm_turbo_rx_lut */
/* replaces srsran_rm_turbo_gentables / srsran_rm_turbo_free_tables (rm_turbo.h:54,56); idempotent, thread-safe */
int  srsb200_rm_turbo_gentables(srsb200_engine_t* e);
/* replaces srsran_rm_turbo_rx_lut / srsran_rm_turbo_rx_lut_ (rm_turbo.h:76-84): output[T[i mod (3K+12)]] += input[i],
 * int16 wrap, natural layout; host buffers; -2 if rv_idx >= 4 or cb_idx >= 188 (rm_turbo.c:441-444) */
int  srsb200_rm_turbo_rx_lut(srsb200_engine_t* e, const int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx,
                             uint32_t rv_idx);
/* the de-rate-matching index table itself (natural layout), 3K+12 entries - for inspection / tests */
int  srsb200_rm_table(uint32_t cb_idx, uint32_t rv_idx, uint16_t* table);

/* ------------------------------------------------------------------ transport blocks: decode_tb / srsran_*sch_decode */
/*
 * One transport-block decode request. Mirrors the arguments of decode_tb (sch.c:509-573) with the soft buffer
 * (lib/include/srsran/phy/fec/softbuffer.h:41-48) passed as its three arrays:
 *   buffer_f[r] -> int16[SRSB200_SOFTBUFFER_SIZE]  accumulating LLR buffer of code block r (in/out, host memory)
 *   sb_data[r]  -> uint8[SRSB200_SOFTBUFFER_SIZE/8] cached bytes of already-decoded code blocks (in/out)
 *   cb_crc[r]   -> per code block CRC flag (in/out); *tb_crc (out)
 * On return: data holds the decoded bytes laid out as the reference does (CB r at byte r*rlen/8), ret is decode_tb's
 * return code (0 ok, -1 CRC failure, -2 invalid inputs), cb_noi[r] the half-iterations run for CB r in this call
 * (0 if skipped) and avg_iterations what srsran_sch_last_noi would report.
 */
typedef struct {
  uint32_t  tbs;
  uint32_t  Qm;          /* bits per symbol x layers, as passed to decode_tb */
  uint32_t  rv;
  uint32_t  nof_e_bits;  /* G */
  const int16_t* e_bits; /* G int16 LLRs */
  int16_t** buffer_f;
  uint8_t** sb_data;
  uint8_t*  cb_crc;      /* bool per CB */
  uint8_t*  tb_crc;
  uint32_t  max_cb;      /* softbuffer->max_cb */
  uint8_t*  data;
  uint32_t* cb_noi;      /* may be NULL */
  float     avg_iterations;
  int       ret;
  /* Optional UL-SCH source (the step right before decode_tb on the eNB path, srsran_ulsch_decode sch.c:1122-1193).
   * When q_bits != NULL, e_bits is ignored: the channel de-interleaver (ulsch_deinterleave sch.c:994-1021 =
   * ulsch_interleave_gen :661-682 + srsran_vec_lut_sis vector.c:147-152) runs on the device,
   *     g[lut[p]] = q[p],  p ascending,  lut[p] = rank of p among the non-RI positions in (row, column, bit) order, 0 for RI,
   * and the transport block is decoded from g + e_offset without leaving the device. */
  const int16_t*  q_bits;        /* H_prime_total*Qm interleaved LLRs (host) */
  uint32_t        H_prime_total; /* nb_q / Qm; must be a multiple of N_pusch_symbs */
  uint32_t        N_pusch_symbs;
  const uint32_t* ri_positions;  /* srsran_uci_bit_t.position of the RI bits (q->ack_ri_bits), may be NULL */
  uint32_t        nof_ri_bits;   /* Q'_ri * Qm */
  uint32_t        e_offset;      /* Q'_cqi * Qm: first g-bit of the UL-SCH data */
  int16_t*        g_bits;        /* optional out (host): first nof_g_out de-interleaved values (CQI decoding reads them) */
  uint32_t        nof_g_out;
  /* Optional descrambling on the device (36.211 6.3.1 / 7.2; srsran_sequence_pdsch_apply_s -> srsran_sequence_apply_s,
   * lib/src/phy/common/sequence.c:507-561): with descramble != 0 the e_bits are the demodulator's output as it is,
   * e_bits[i] is negated where the Gold sequence of seed c_init has c(i) = 1 (int16 wrap), fused into the rate de-matcher.
   * c_init = rnti*2^14 + q*2^13 + floor(ns/2)*2^9 + cell_id for the PDSCH. Not combined with the q_bits source (on the
   * uplink the host-side RI/ACK decoding needs the descrambled values first). */
  uint32_t        descramble;
  uint32_t        c_init;
  /* Half-iteration limit of THIS transport block (q->max_iterations of its srsran_sch_t, sch.c:223-230); 0 = the
   * max_iterations argument of the call. Blocks of one submission may differ. */
  uint32_t        max_iterations;
  /* q->llr_is_8bit (lib/include/srsran/phy/phch/sch.h:57): e_bits points to int8 LLRs and buffer_f[r] to int8[SRSB200_SOFTBUFFER_SIZE]
   * (the reference casts the same arrays, sch.c:410,428). Rate de-matching is srsran_rm_turbo_rx_lut_8bit (wrapping int8), the
   * decoder what srsran_tdec_iteration_8bit runs in AUTO mode - the windowed saturating int8 decoders - bit for bit. Accepted
   * for transport blocks whose code-block sizes have an 8-bit decoder in the reference (srsb200_tdec8_windows(K) != 0: K > 800
   * and K % 16 == 0; the reference widens smaller blocks to its SSE int16 decoders, the drop-in leaves those to the reference's
   * own loop), without q_bits / descramble. All transport blocks of one submission must agree on this flag. */
  uint32_t        llr_is_8bit;
  /* Optional modulation-symbol source: with symbols != NULL the soft demodulator runs on the device -
   * srsran_demod_soft_demodulate_s (lib/src/phy/modem/demod_soft.c:871-894; called at pdsch.c:696 and pusch.c:422) bit for bit,
   * including which symbols the reference's SSE / AVX2 bodies round and which its scalar tails truncate - and e_bits (downlink)
   * or q_bits (uplink: H_prime_total != 0) are ignored: the host sends 8 bytes per resource element instead of 2 * Qm.
   * symbols = nof_symbols (re, im) float pairs of equalised symbols (q->d), mod = srsran_mod_t (0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM,
   * 4 256QAM). Usually combined with descramble / c_init: downlink as above; uplink the interleaved stream is descrambled as
   * a whole (srsran_sequence_pusch_apply_s, pusch.c:439-441) before the de-interleaver, and the host-side UCI decoding gets the
   * descrambled LLRs it reads back through q_gather: q_gather_out[i] = q[q_gather_pos[i]] (RI / ACK positions). int16 mode only. */
  const float*    symbols;
  uint32_t        nof_symbols;
  uint32_t        mod;
  const uint32_t* q_gather_pos;
  int16_t*        q_gather_out;
  uint32_t        nof_q_gather;
} srsb200_tb_t;

/*
 * HARQ soft buffers (srsran_softbuffer_rx_t, lib/src/phy/fec/softbuffer.c:36-178). Default = host-coherent: every
 * decode_tb call uploads buffer_f[r] of the code blocks it decodes and writes the combined LLRs back, so the caller's
 * memory is always what the reference would hold. With srsb200_softbuffer_set_resident(e, 1) the engine keeps a device
 * mirror keyed by the host pointer buffer_f[r] instead: no per-call copies; the caller must then forward the
 * reference's reset points (srsran_softbuffer_rx_reset / _reset_tbs / _reset_cb, softbuffer.c:139-169) to
 * srsb200_softbuffer_reset, may fetch the contents with srsb200_softbuffer_sync_to_host, and frees the mirror in
 * srsran_softbuffer_rx_free via srsb200_softbuffer_release. A buffer first seen without a reset adopts the host content.
 * srsb200_softbuffer_reset costs no launch of its own: the mirrors it names are zeroed by one kernel in front of the next call that
 * reads or writes a mirror (a transport-block submission, sync_to_host), so a reset per transport block and subframe is cheap.
 */
int srsb200_softbuffer_set_resident(srsb200_engine_t* e, int resident);
int srsb200_softbuffer_reset(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb);
int srsb200_softbuffer_sync_to_host(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb);
int srsb200_softbuffer_release(srsb200_engine_t* e, int16_t** buffer_f, uint32_t nof_cb);

/* replaces srsran_demod_soft_demodulate_s (lib/src/phy/modem/demod_soft.c:871-894), host in / host out: llr receives
 * nsymbols * bits-per-symbol values; -1 for an unknown modulation (like the reference) */
int srsb200_demod_soft_demodulate_s(srsb200_engine_t* e, uint32_t mod, const float* symbols, int16_t* llr, uint32_t nsymbols);

/* UL-SCH channel de-interleaver alone (host in, host out): g_bits receives H_prime_total*Qm values, of which the first
 * H_prime_total*Qm - nof_ri_bits are defined (as in the reference) */
int srsb200_ulsch_deinterleave(srsb200_engine_t* e, const int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                               int16_t* g_bits, const uint32_t* ri_positions, uint32_t nof_ri_bits);

/* decode n transport blocks (possibly of different UEs/cells/subframes) as one batched submission */
int srsb200_decode_tb_batch(srsb200_engine_t* e, srsb200_tb_t* tbs, uint32_t n, uint32_t max_iterations);
/* single transport block = decode_tb (sch.c:509); returns tb->ret */
int srsb200_decode_tb(srsb200_engine_t* e, srsb200_tb_t* tb, uint32_t max_iterations);

/* ------------------------------------------------------------------------------------------------------------------
 * Transport-block ENCODE (the transmit mirror of the path)
 * replaces: static encode_tb_off / encode_tb, lib/src/phy/phch/sch.c:240-375 (hence srsran_dlsch_encode[2], sch.c:611-639,
 *           and the data part of srsran_ulsch_encode), i.e. per code block
 *           srsran_tcod_encode_lut  lib/src/phy/fec/turbo/turbocoder.c:186-360  (CRC24A/CRC24B attach + turbo encode)
 *           srsran_rm_turbo_tx_lut  lib/src/phy/fec/turbo/rm_turbo.c:345-388    (circular buffer + bit selection)
 * Semantics kept: K2 (smaller) code blocks first (sch.c:287-293), E = Qm*floor(G'/C) for r <= C-gamma-1 else
 * Qm*ceil(G'/C) (sch.c:299-303), code block r packed MSB-first at bit offset sum(E_0..E_{r-1}); return codes: -2 for
 * null pointers, -1 for filler bits / Qm == 0 / C > max_cb / rv > 3 / segmentation failure, 0 otherwise (tbs == 0 writes
 * nothing). Differences: every call encodes from `data` (the reference re-reads its circular buffer softbuffer->buffer_b
 * for rv != 0 and data == NULL is accepted there; here data == NULL is -2), and the (nof_e_bits+7)/8 output bytes are
 * written whole (pad bits zero) instead of bit-spliced into the previous content.
 */
typedef struct {
  uint32_t       tbs;        /* cb_segm->tbs */
  uint32_t       Qm;
  uint32_t       rv;
  uint32_t       nof_e_bits; /* G */
  uint32_t       max_cb;     /* softbuffer->max_cb */
  const uint8_t* data;       /* tbs/8 payload bytes (host) */
  uint8_t*       e_bits;     /* out (host): packed, (nof_e_bits+7)/8 bytes */
  int32_t        ret;        /* out: per-TB return code */
} srsb200_tb_tx_t;
int srsb200_encode_tb_batch(srsb200_engine_t* e, srsb200_tb_tx_t* tbs, uint32_t n);
int srsb200_encode_tb(srsb200_engine_t* e, srsb200_tb_tx_t* tb); /* returns tb->ret */

/* ------------------------------------------------------------------------------------------------------------------
 * Several GPUs in one process: independent units, no exchange step, hence no collective - the dispatcher is
 * host-side. One engine + one host thread per device; a submission is split by owner and the parts run concurrently.
 * Reference analogue: one srsran_sch_t per PHY worker thread and carrier (srsenb/src/phy/lte/worker_pool.cc:32-58,
 * cc_worker.cc:345). `owner[i]` = any id that is stable per HARQ entity (cell id, cell * nof_ue + ue, ...): the transport
 * block goes to device owner[i] mod nof_devices, every time - soft-buffer mirrors stay where they are. owner == NULL: the
 * soft buffer's address is the key.
 */
typedef struct srsb200_multi srsb200_multi_t;
int  srsb200_device_count(void);
/* devices == NULL / n == 0: every CUDA device of the process */
int  srsb200_multi_create(srsb200_multi_t** m, const int* devices, int n);
void srsb200_multi_destroy(srsb200_multi_t* m);
int  srsb200_multi_nof_devices(const srsb200_multi_t* m);
srsb200_engine_t* srsb200_multi_engine(srsb200_multi_t* m, int i);           /* e.g. for srsb200_softbuffer_set_resident per device */
int  srsb200_multi_device_of(const srsb200_multi_t* m, uint64_t owner);      /* index into the device list */
int  srsb200_multi_decode_tb_batch(srsb200_multi_t* m, srsb200_tb_t* tbs, uint32_t n, const uint64_t* owner, uint32_t max_iterations);
/* flat code-block batch (srsb200_tdec_batch's arguments): contiguous ranges balanced by sum(K), one per device */
int  srsb200_multi_tdec_batch(srsb200_multi_t* m, uint32_t n, const uint32_t* K, const uint8_t* crc_kind, const int16_t* llr,
                              const uint64_t* llr_offset, uint64_t llr_len, uint32_t max_iter, uint32_t min_iter, int early_stop,
                              uint8_t* out_bytes, const uint64_t* out_offset, uint64_t out_len, uint8_t* noi, uint8_t* crc_ok);

#ifdef __cplusplus
}
#endif
#endif /* SRSRAN_B200_H */
