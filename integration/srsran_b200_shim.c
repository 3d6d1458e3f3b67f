/*
 * srsran_b200_shim.c - the reference-side binding: srsRAN 4G's own symbols for the turbo-decode hot path, implemented on
 * top of the C ABI of libsrsran_b200.so (include/srsran_b200.h).
 *
 * This file is compiled INSIDE the reference tree (it includes the reference's own headers for the struct layouts, so
 * sizeof(srsran_tdec_t) / sizeof(srsran_sch_t) stay what already-compiled callers expect) and replaces the objects
 *   lib/src/phy/fec/turbo/turbodecoder.c      (srsran_tdec_*)
 *   the rx half of lib/src/phy/fec/turbo/rm_turbo.c (srsran_rm_turbo_rx_lut[_])
 * and provides srsran_b200_decode_tb(), which lib/src/phy/phch/sch.c:decode_tb calls instead of its serial
 * decode_tb_cb loop (two-line patch), and srsran_b200_encode_tb() for sch.c:encode_tb_off likewise. Everything above decode_tb (srsran_dlsch_decode[2],
 * srsran_ulsch_decode with its UCI de-multiplexing) is untouched and funnels into it exactly as before.
 *
 * The device handle lives in fields the reference already has: h->dec16_hdlr[0] (srsb200_tdec_t*). One engine per
 * calling thread and device, created on first use, destroyed at thread exit (devices: $SRSRAN_B200_DEVICES / $SRSRAN_B200_DEVICE,
 * below); the C ABI serialises calls per engine, so any number of PHY worker threads may call concurrently, like the lock-free
 * reference objects, and their submissions overlap.
 *
 * Layout contract: srsran_tdec_autoimp_get_subblocks() returns 0 for every size, which makes srsran_rm_turbo_rx_lut and
 * the decoder agree on the natural (generic decoder) input order - SURVEY.md section 8(b).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "srsran/phy/fec/cbsegm.h"
#include "srsran/phy/fec/softbuffer.h"
#include "srsran/phy/fec/turbo/rm_turbo.h"
#include "srsran/phy/fec/turbo/turbodecoder.h"
#include "srsran/phy/phch/sch.h"
#include "srsran/phy/phch/uci_cfg.h"
#include "srsran/phy/utils/debug.h"

#include "srsran_b200.h"

/*
 * Engines: one per (calling thread, device), created on first use and destroyed when the thread exits (pthread key destructor).
 * srsRAN runs one srsran_sch_t per PHY worker thread (x carrier) with no locks between them, and a decode is latency-bound
 * (sequential recursions), so the submissions of different workers must be able to overlap on the GPU - which they do on
 * separate engines (own streams and workspaces; examples/multicell_uplink.c measures it). Decoder objects remember the engine
 * that created them and stay usable from any thread.
 *
 * Devices: SRSRAN_B200_DEVICES = "all" or a comma-separated list ("0,1,2,3") turns the shim into a multi-GPU dispatcher for a
 * multi-cell eNB process (reference analogue: srsenb/src/phy/lte/worker_pool.cc:32-58 - one srsran_sch_t per worker and
 * carrier); otherwise the single device SRSRAN_B200_DEVICE (default 0). Placement is by OWNER KEY modulo the device count: the
 * soft buffer's address for transport blocks (a HARQ process always lands on the same GPU: SURVEY.md 8(e) "HARQ state is
 * sticky"), the decoder object's address for srsran_tdec_*. No data crosses between devices - code blocks are independent.
 */
#define SHIM_MAX_DEV 16
typedef struct {
  srsb200_engine_t* eng[SHIM_MAX_DEV];
} thread_engines_t;

static pthread_key_t  g_key;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int            g_ndev = 1;
static int            g_dev[SHIM_MAX_DEV];

static void thread_engines_free(void* p)
{
  thread_engines_t* te = (thread_engines_t*)p;
  if (te) {
    for (int i = 0; i < SHIM_MAX_DEV; i++) {
      if (te->eng[i]) {
        srsb200_engine_destroy(te->eng[i]); /* streams, pinned arenas, device tables and scratch of this thread */
      }
    }
    free(te);
  }
}

static void shim_init(void)
{
  pthread_key_create(&g_key, thread_engines_free);
  const char* list = getenv("SRSRAN_B200_DEVICES");
  const char* one  = getenv("SRSRAN_B200_DEVICE");
  g_ndev   = 1;
  g_dev[0] = one ? atoi(one) : 0;
  if (list && *list) {
    int n = 0;
    if (!strcasecmp(list, "all")) {
      int total = srsb200_device_count();
      for (int d = 0; d < total && n < SHIM_MAX_DEV; d++) {
        g_dev[n++] = d;
      }
    } else {
      for (const char* p = list; *p && n < SHIM_MAX_DEV;) {
        g_dev[n++] = atoi(p);
        while (*p && *p != ',') {
          p++;
        }
        if (*p == ',') {
          p++;
        }
      }
    }
    if (n > 0) {
      g_ndev = n;
    }
  }
}

static srsb200_engine_t* engine_slot(int slot)
{
  pthread_once(&g_once, shim_init);
  thread_engines_t* te = (thread_engines_t*)pthread_getspecific(g_key);
  if (te == NULL) {
    te = calloc(1, sizeof(thread_engines_t));
    if (te == NULL || pthread_setspecific(g_key, te)) {
      free(te);
      return NULL;
    }
  }
  if (te->eng[slot] == NULL) {
    if (srsb200_engine_create(&te->eng[slot], g_dev[slot]) != SRSB200_SUCCESS) {
      ERROR("srsran_b200: %s", srsb200_last_error());
      te->eng[slot] = NULL;
    }
  }
  return te->eng[slot];
}

/* the calling thread's engine on the device that owns `key` */
static srsb200_engine_t* engine_for(const void* key)
{
  pthread_once(&g_once, shim_init);
  uint64_t k = (uint64_t)(uintptr_t)key;
  k ^= k >> 17; /* allocator addresses share their low and high bits: mix before the modulo */
  k *= 0x9E3779B97F4A7C15ull;
  return engine_slot((int)((k >> 32) % (uint64_t)g_ndev));
}

static srsb200_engine_t* engine(void)
{
  return engine_slot(0);
}

/* which device (index into the list) serves this soft buffer / decoder object - for tests and operators */
int srsran_b200_device_of(const void* key)
{
  pthread_once(&g_once, shim_init);
  uint64_t k = (uint64_t)(uintptr_t)key;
  k ^= k >> 17;
  k *= 0x9E3779B97F4A7C15ull;
  return g_dev[(k >> 32) % (uint64_t)g_ndev];
}
int srsran_b200_nof_devices(void)
{
  pthread_once(&g_once, shim_init);
  return g_ndev;
}

/* ------------------------------------------------------------------ srsran_tdec_* (turbodecoder.h:97-116) */
int srsran_tdec_init(srsran_tdec_t* h, uint32_t max_long_cb)
{
  return srsran_tdec_init_manual(h, max_long_cb, SRSRAN_TDEC_AUTO);
}

int srsran_tdec_init_manual(srsran_tdec_t* h, uint32_t max_long_cb, srsran_tdec_impl_type_t dec_type)
{
  bzero(h, sizeof(srsran_tdec_t));
  h->dec_type    = dec_type; /* every implementation type maps to the one bit-exact device decoder */
  h->max_long_cb = max_long_cb;
  srsb200_tdec_t* d = NULL;
  if (srsb200_tdec_init(&d, engine_for(h), max_long_cb) != SRSB200_SUCCESS) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  h->dec16_hdlr[0]    = d;
  h->current_llr_type = SRSRAN_TDEC_16;
  h->current_cbidx    = -1;
  return SRSRAN_SUCCESS;
}

void srsran_tdec_free(srsran_tdec_t* h)
{
  if (h->dec16_hdlr[0]) {
    srsb200_tdec_free((srsb200_tdec_t*)h->dec16_hdlr[0]);
  }
  bzero(h, sizeof(srsran_tdec_t));
}

void srsran_tdec_force_not_sb(srsran_tdec_t* h)
{
  h->force_not_sb = true; /* the device decoder always takes the natural layout */
}

int srsran_tdec_new_cb(srsran_tdec_t* h, uint32_t long_cb)
{
  if (srsb200_tdec_new_cb((srsb200_tdec_t*)h->dec16_hdlr[0], long_cb) != SRSB200_SUCCESS) {
    ERROR("%s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  h->n_iter          = 0;
  h->current_long_cb = long_cb;
  h->current_cbidx   = srsran_cbsegm_cbindex(long_cb);
  return SRSRAN_SUCCESS;
}

int srsran_tdec_get_nof_iterations(srsran_tdec_t* h)
{
  return h->n_iter;
}

uint32_t srsran_tdec_autoimp_get_subblocks(uint32_t long_cb)
{
  return srsb200_tdec_autoimp_get_subblocks(long_cb);
}

uint32_t srsran_tdec_autoimp_get_subblocks_8bit(uint32_t long_cb)
{
  return srsb200_tdec_autoimp_get_subblocks(long_cb);
}

void srsran_tdec_iteration(srsran_tdec_t* h, int16_t* input, uint8_t* output)
{
  if (h->current_cbidx < 0) {
    ERROR("Error CB index not set (call srsran_tdec_new_cb() first");
    return;
  }
  if (srsb200_tdec_iteration((srsb200_tdec_t*)h->dec16_hdlr[0], input, output) != SRSB200_SUCCESS) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return;
  }
  h->n_iter++;
}

int srsran_tdec_run_all(srsran_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  if (srsran_tdec_new_cb(h, long_cb)) {
    return SRSRAN_ERROR;
  }
  if (srsb200_tdec_run_all((srsb200_tdec_t*)h->dec16_hdlr[0], input, output, nof_iterations, long_cb) != SRSB200_SUCCESS) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  h->n_iter = srsb200_tdec_get_nof_iterations((srsb200_tdec_t*)h->dec16_hdlr[0]);
  return SRSRAN_SUCCESS;
}

/*
 * 8-bit LLR entry points (turbodecoder.h:118-121). The reference's 8-bit mode runs its windowed saturating int8 decoders
 * (turbodecoder_win.h with llr_t = int8_t) for K > 800 with K % 16 == 0; the engine reproduces those bit for bit
 * (srsb200_tdec_batch8). For the smaller sizes the reference itself leaves 8-bit arithmetic (it widens to its SSE int16
 * decoders, turbodecoder.c:443-476); there the LLRs are widened and decoded by the exact int16 engine - same API, results of the
 * generic decoder on those values. No allocation per call: one scratch buffer per thread.
 */
static __thread int16_t t_widen[3 * 6144 + 12];
static int tdec8_run(srsran_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  uint32_t K = long_cb, kind = 0;
  uint64_t loff = 0, ooff = 0;
  uint8_t  noi = 0, ok = 0;
  uint8_t  kind8 = (uint8_t)kind;
  /* stateless: half-iteration n of the 8-bit decoder is a pure function of the input, so "one more half-iteration" is the decode
   * run to n + 1 from the start (the transport-block path, where it matters, submits whole decodes through the sch.c hook) */
  int r = srsb200_tdec_batch8(engine_for(h), 1, &K, &kind8, input, &loff, 3ull * K + 12, nof_iterations ? nof_iterations : 1, 1, 0, output, &ooff, K / 8,
                              &noi, &ok);
  if (r != SRSB200_SUCCESS) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  return SRSRAN_SUCCESS;
}
void srsran_tdec_iteration_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output)
{
  if (h->current_cbidx < 0) {
    ERROR("Error CB index not set (call srsran_tdec_new_cb() first");
    return;
  }
  if (srsb200_tdec8_windows(h->current_long_cb)) {
    if (tdec8_run(h, input, output, (uint32_t)h->n_iter + 1, h->current_long_cb) == SRSRAN_SUCCESS) {
      h->n_iter++;
    }
    return;
  }
  uint32_t n = 3 * h->current_long_cb + 12;
  for (uint32_t i = 0; i < n; i++) {
    t_widen[i] = input[i];
  }
  srsran_tdec_iteration(h, t_widen, output);
}
int srsran_tdec_run_all_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  if (srsb200_tdec8_windows(long_cb)) {
    if (srsran_tdec_new_cb(h, long_cb)) {
      return SRSRAN_ERROR;
    }
    int ret = tdec8_run(h, input, output, nof_iterations, long_cb);
    if (ret == SRSRAN_SUCCESS) {
      h->n_iter = nof_iterations ? (int)nof_iterations : 1;
    }
    return ret;
  }
  if (long_cb > 6144) {
    return SRSRAN_ERROR;
  }
  uint32_t n = 3 * long_cb + 12;
  for (uint32_t i = 0; i < n; i++) {
    t_widen[i] = input[i];
  }
  return srsran_tdec_run_all(h, t_widen, output, nof_iterations, long_cb);
}

/* Names the reference does not have but integrators ask for: thin aliases. */
int srsran_tdec_get_hard_decision(srsran_tdec_t* h, uint8_t* output, uint32_t long_cb)
{
  if (h == NULL || output == NULL || long_cb != h->current_long_cb) {
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  return srsb200_tdec_get_hard_decision((srsb200_tdec_t*)h->dec16_hdlr[0], output) == SRSB200_SUCCESS ? SRSRAN_SUCCESS : SRSRAN_ERROR;
}

/* ------------------------------------------------------------------ srsran_rm_turbo_rx_lut (rm_turbo.h:54-84), rx half */
void srsran_b200_rm_turbo_gentables(void)
{
  srsb200_rm_turbo_gentables(engine());
}

int srsran_rm_turbo_rx_lut_(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx, bool enable_input_tdec)
{
  (void)enable_input_tdec; /* natural layout in both cases (autoimp_get_subblocks == 0) */
  return srsb200_rm_turbo_rx_lut(engine(), input, output, in_len, cb_idx, rv_idx);
}

int srsran_rm_turbo_rx_lut(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx)
{
  return srsran_rm_turbo_rx_lut_(input, output, in_len, cb_idx, rv_idx, true);
}

/* rm_turbo.h:86-87: wrapping int8 accumulation, natural layout (srsran_tdec_autoimp_get_subblocks_8bit reports 0 here) */
int srsran_rm_turbo_rx_lut_8bit(int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx)
{
  return srsb200_rm_turbo_rx_lut8(engine(), input, output, in_len, cb_idx, rv_idx);
}

/* Does the batched entry point take this transport block? int16 LLRs: always. q->llr_is_8bit: when its code-block sizes run in the
 * reference's 8-bit window decoders (K > 800, K % 16 == 0) - smaller blocks stay in the reference's own loop (sch.c:371-494), which
 * then calls the per-block 8-bit symbols above. */
int srsran_b200_takes_tb(srsran_sch_t* q, srsran_cbsegm_t* cb_segm)
{
  if (q == NULL || cb_segm == NULL || !q->llr_is_8bit) {
    return 1;
  }
  if (cb_segm->C == 0) {
    return 1;
  }
  return srsb200_tdec8_windows(cb_segm->K1) != 0 && (cb_segm->C2 == 0 || srsb200_tdec8_windows(cb_segm->K2) != 0);
}

/* ------------------------------------------------------------------ decode_tb (sch.c:509-573) */
/*
 * Drop-in body of sch.c's static decode_tb(): same arguments, same return codes (0 ok, -1 CRC failure, -2 invalid
 * inputs), same side effects on the soft buffer (buffer_f accumulation, cb_crc / tb_crc, cached bytes) and on
 * q->avg_iterations. All code blocks of the transport block go to the device as one batched submission.
 */
static int decode_tb_common(srsran_sch_t* q, srsran_softbuffer_rx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv,
                            uint32_t nof_e_bits, int16_t* e_bits, uint8_t* data, int descramble, uint32_t c_init);

int srsran_b200_decode_tb(srsran_sch_t*           q,
                          srsran_softbuffer_rx_t* softbuffer,
                          srsran_cbsegm_t*        cb_segm,
                          uint32_t                Qm,
                          uint32_t                rv,
                          uint32_t                nof_e_bits,
                          int16_t*                e_bits,
                          uint8_t*                data)
{
  return decode_tb_common(q, softbuffer, cb_segm, Qm, rv, nof_e_bits, e_bits, data, 0, 0);
}

/*
 * The same with the descrambling of srsran_pdsch_decode (pdsch.c:726-732, srsran_sequence_pdsch_apply_s) moved onto the
 * device: e_bits is the soft demodulator's output as it is, c_init the scrambling seed of the codeword
 * (rnti << 14 | codeword << 13 | (nslot / 2) << 9 | cell id, 36.211 6.3.1). A caller in pdsch.c skips its own
 * srsran_sequence_pdsch_apply_s and hands the seed down instead.
 */
int srsran_b200_decode_tb_scrambled(srsran_sch_t*           q,
                                    srsran_softbuffer_rx_t* softbuffer,
                                    srsran_cbsegm_t*        cb_segm,
                                    uint32_t                Qm,
                                    uint32_t                rv,
                                    uint32_t                nof_e_bits,
                                    int16_t*                e_bits,
                                    uint8_t*                data,
                                    uint32_t                c_init)
{
  return decode_tb_common(q, softbuffer, cb_segm, Qm, rv, nof_e_bits, e_bits, data, 1, c_init);
}

static int decode_tb_common(srsran_sch_t* q, srsran_softbuffer_rx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv,
                            uint32_t nof_e_bits, int16_t* e_bits, uint8_t* data, int descramble, uint32_t c_init)
{
  if (q == NULL || data == NULL || softbuffer == NULL || e_bits == NULL || cb_segm == NULL || Qm == 0) {
    ERROR("Missing inputs: data=%d, softbuffer=%d, e_bits=%d, cb_segm=%d Qm=%d", data != 0, softbuffer != 0, e_bits != 0, cb_segm != 0, Qm);
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  if (q->llr_is_8bit && (descramble || !srsran_b200_takes_tb(q, cb_segm))) {
    ERROR("srsran_b200: this 8-bit transport block stays in the reference's per-code-block loop (srsran_b200_takes_tb)");
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint8_t      tb_crc = 0;
  srsb200_tb_t tb;
  memset(&tb, 0, sizeof(tb));
  tb.tbs        = cb_segm->tbs;
  tb.Qm         = Qm;
  tb.rv         = rv;
  tb.nof_e_bits = nof_e_bits;
  tb.e_bits     = e_bits;
  tb.buffer_f   = softbuffer->buffer_f;
  tb.sb_data    = softbuffer->data;
  tb.cb_crc     = (uint8_t*)softbuffer->cb_crc; /* bool[] */
  tb.tb_crc     = &tb_crc;
  tb.max_cb     = softbuffer->max_cb;
  tb.data       = data;
  tb.cb_noi     = NULL;
  tb.descramble  = descramble ? 1u : 0u;
  tb.c_init      = c_init;
  tb.llr_is_8bit = q->llr_is_8bit ? 1u : 0u; /* e_bits / buffer_f then point to int8 data, exactly as sch.c:410,428 casts them */
  int ret = srsb200_decode_tb(engine_for(softbuffer), &tb, q->max_iterations);
  if (ret == SRSB200_ERROR_NO_DEVICE) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  if (cb_segm->tbs != 0 && cb_segm->C != 0 && ret != SRSRAN_ERROR_INVALID_INPUTS) {
    softbuffer->tb_crc = tb_crc != 0;
    q->avg_iterations  = tb.avg_iterations;
  }
  return ret;
}

/*
 * The whole receive tail of srsran_pdsch_decode for one codeword (pdsch.c:693-740) in one device submission: the equalised symbols
 * q->d[cw] go to the device as they are (8 bytes per resource element instead of 2 * Qm bytes of LLRs), soft demodulation
 * (srsran_demod_soft_demodulate_s), descrambling (srsran_sequence_pdsch_apply_s) and decode_tb run there. A caller in pdsch.c
 * replaces its three calls by this one; mod is the codeword's srsran_mod_t, Qm what srsran_dlsch_decode2 would pass to decode_tb
 * (bits per symbol x Nl), c_init the scrambling seed as for srsran_b200_decode_tb_scrambled.
 */
int srsran_b200_decode_tb_symbols(srsran_sch_t*           q,
                                  srsran_softbuffer_rx_t* softbuffer,
                                  srsran_cbsegm_t*        cb_segm,
                                  srsran_mod_t            mod,
                                  uint32_t                Qm,
                                  uint32_t                rv,
                                  uint32_t                nof_e_bits,
                                  cf_t*                   symbols,
                                  uint32_t                nof_symbols,
                                  uint32_t                c_init,
                                  uint8_t*                data)
{
  if (q == NULL || data == NULL || softbuffer == NULL || symbols == NULL || cb_segm == NULL || Qm == 0 || q->llr_is_8bit) {
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint8_t      tb_crc = 0;
  srsb200_tb_t tb;
  memset(&tb, 0, sizeof(tb));
  tb.tbs         = cb_segm->tbs;
  tb.Qm          = Qm;
  tb.rv          = rv;
  tb.nof_e_bits  = nof_e_bits;
  tb.buffer_f    = softbuffer->buffer_f;
  tb.sb_data     = softbuffer->data;
  tb.cb_crc      = (uint8_t*)softbuffer->cb_crc;
  tb.tb_crc      = &tb_crc;
  tb.max_cb      = softbuffer->max_cb;
  tb.data        = data;
  tb.symbols     = (const float*)symbols;
  tb.nof_symbols = nof_symbols;
  tb.mod         = (uint32_t)mod;
  tb.descramble  = 1;
  tb.c_init      = c_init;
  int ret = srsb200_decode_tb(engine_for(softbuffer), &tb, q->max_iterations);
  if (ret == SRSB200_ERROR_NO_DEVICE) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  if (cb_segm->tbs != 0 && cb_segm->C != 0 && ret != SRSRAN_ERROR_INVALID_INPUTS) {
    softbuffer->tb_crc = tb_crc != 0;
    q->avg_iterations  = tb.avg_iterations;
  }
  return ret;
}

/*
 * What srsran_pdsch_codeword_decode (pdsch.c:662-760) calls under -DSRSRAN_B200 instead of srsran_demod_soft_demodulate_s (:696),
 * srsran_sequence_pdsch_apply_s (:726-732) and srsran_dlsch_decode2 (:740): the arguments of srsran_dlsch_decode2 (sch.c:580-609)
 * with the equalised symbols of the codeword in place of its LLRs, plus the four values the scrambling seed is made of
 * (sequence_pdsch_seed, sequences.c:62-65). Same return codes as srsran_dlsch_decode2.
 */
uint32_t srsran_b200_pdsch_c_init(uint16_t rnti, int codeword_idx, uint32_t nslot, uint32_t cell_id)
{
  return ((uint32_t)rnti << 14) + ((uint32_t)codeword_idx << 13) + ((nslot / 2) << 9) + cell_id;
}

/* 1: this codeword can go symbols-in (16-bit LLR mode, a transport block the engine decodes); 0: keep the reference's three calls */
int srsran_b200_dlsch_takes_symbols(srsran_sch_t* q, srsran_pdsch_cfg_t* cfg, int tb_idx)
{
  if (q == NULL || cfg == NULL || tb_idx < 0 || tb_idx >= SRSRAN_MAX_CODEWORDS || q->llr_is_8bit) {
    return 0;
  }
  if (cfg->softbuffers.rx[tb_idx] == NULL || cfg->grant.tb[tb_idx].tbs <= 0 || cfg->grant.nof_re == 0) {
    return 0;
  }
  uint32_t sg[8];
  return srsb200_cbsegm((uint32_t)cfg->grant.tb[tb_idx].tbs, sg) == SRSB200_SUCCESS && sg[0] == 0; /* no filler bits */
}

int srsran_b200_dlsch_decode2_symbols(srsran_sch_t*       q,
                                      srsran_pdsch_cfg_t* cfg,
                                      cf_t*               symbols,
                                      uint32_t            c_init,
                                      uint8_t*            data,
                                      int                 tb_idx,
                                      uint32_t            nof_layers)
{
  if (q == NULL || cfg == NULL || symbols == NULL || data == NULL || tb_idx < 0 || tb_idx >= SRSRAN_MAX_CODEWORDS) {
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint32_t Nl = (nof_layers != cfg->grant.nof_tb) ? 2 : 1; /* sch.c:587-591 */
  uint32_t        sg[8];
  srsran_cbsegm_t cb_segm;
  memset(&cb_segm, 0, sizeof(cb_segm));
  if (srsb200_cbsegm((uint32_t)cfg->grant.tb[tb_idx].tbs, sg) != SRSB200_SUCCESS) {
    ERROR("Error computing Codeword (%d) segmentation for TBS=%d", tb_idx, cfg->grant.tb[tb_idx].tbs);
    return SRSRAN_ERROR;
  }
  cb_segm.F = sg[0]; cb_segm.C = sg[1]; cb_segm.K1 = sg[2]; cb_segm.K2 = sg[3]; cb_segm.K1_idx = sg[4]; cb_segm.K2_idx = sg[5];
  cb_segm.C1 = sg[6]; cb_segm.C2 = sg[7];
  cb_segm.tbs = (uint32_t)cfg->grant.tb[tb_idx].tbs; cb_segm.L_tb = 24; cb_segm.L_cb = 24;
  static const uint32_t bits_x_symbol[5] = {1, 2, 4, 6, 8}; /* srsran_mod_bits_x_symbol (phy_common.c): BPSK .. 256QAM */
  if ((uint32_t)cfg->grant.tb[tb_idx].mod > 4) {
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint32_t Qm = bits_x_symbol[cfg->grant.tb[tb_idx].mod];
  /* the reference demodulates cfg->grant.nof_re symbols of q->d[codeword] whatever the number of layers (pdsch.c:696: layer
   * de-mapping has already gathered the codeword's symbols, :885-887) and hands decode_tb Qm x Nl (sch.c:601-608) */
  return srsran_b200_decode_tb_symbols(q, cfg->softbuffers.rx[tb_idx], &cb_segm, cfg->grant.tb[tb_idx].mod, Qm * Nl, (uint32_t)cfg->grant.tb[tb_idx].rv,
                                       cfg->grant.tb[tb_idx].nof_bits, symbols, cfg->grant.nof_re, c_init, data);
}

/* srsran_demod_soft_demodulate_s (demod_soft.h) on the device, same return codes */
int srsran_b200_demod_soft_demodulate_s(srsran_mod_t modulation, const cf_t* symbols, short* llr, int nsymbols)
{
  int r = srsb200_demod_soft_demodulate_s(engine(), (uint32_t)modulation, (const float*)symbols, llr, nsymbols < 0 ? 0u : (uint32_t)nsymbols);
  return r == SRSB200_SUCCESS ? 0 : -1;
}

int srsran_sch_decode(srsran_sch_t* q, srsran_softbuffer_rx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits,
                      int16_t* e_bits, uint8_t* data)
{
  return srsran_b200_decode_tb(q, softbuffer, cb_segm, Qm, rv, nof_e_bits, e_bits, data);
}

/*
 * Batched entry point for callers that hold several transport blocks at once (both codewords of a PDSCH, all PUSCH
 * grants of a subframe, several cells): one device submission for all of them. results[i] receives decode_tb's code.
 */
int srsran_b200_decode_tb_batch(srsran_sch_t**           q,
                                srsran_softbuffer_rx_t** softbuffer,
                                srsran_cbsegm_t*         cb_segm,
                                const uint32_t*          Qm,
                                const uint32_t*          rv,
                                const uint32_t*          nof_e_bits,
                                int16_t**                e_bits,
                                uint8_t**                data,
                                uint32_t                 n,
                                int*                     results)
{
  srsb200_tb_t* tb     = calloc(n ? n : 1, sizeof(srsb200_tb_t));
  uint8_t*      tb_crc = calloc(n ? n : 1, 1);
  uint32_t*     order  = calloc(n ? n : 1, sizeof(uint32_t));
  if (!tb || !tb_crc || !order) {
    free(tb);
    free(tb_crc);
    free(order);
    for (uint32_t i = 0; i < n; i++) {
      results[i] = SRSRAN_ERROR;
    }
    return SRSRAN_ERROR;
  }
  /* group the transport blocks by owning device (stable within a device), one submission per device */
  uint32_t m = 0;
  int      ret = SRSB200_SUCCESS;
  for (int d = 0; d < srsran_b200_nof_devices(); d++) {
    uint32_t first = m;
    srsb200_engine_t* e = NULL;
    for (uint32_t i = 0; i < n; i++) {
      srsb200_engine_t* ei = engine_for(softbuffer[i]);
      if (srsran_b200_device_of(softbuffer[i]) != g_dev[d]) {
        continue;
      }
      e                     = ei;
      order[m]              = i;
      tb[m].tbs             = cb_segm[i].tbs;
      tb[m].Qm              = Qm[i];
      tb[m].rv              = rv[i];
      tb[m].nof_e_bits      = nof_e_bits[i];
      tb[m].e_bits          = e_bits[i];
      tb[m].buffer_f        = softbuffer[i]->buffer_f;
      tb[m].sb_data         = softbuffer[i]->data;
      tb[m].cb_crc          = (uint8_t*)softbuffer[i]->cb_crc;
      tb[m].tb_crc          = &tb_crc[i];
      tb[m].max_cb          = softbuffer[i]->max_cb;
      tb[m].data            = data[i];
      tb[m].max_iterations  = q[i]->max_iterations; /* every transport block stops at the limit of ITS srsran_sch_t */
      m++;
    }
    if (m > first) {
      /* (devices are served one after the other from this thread; callers that want the devices to overlap submit from one
       *  thread per cell, as srsENB's worker pool does, or use srsb200_multi_decode_tb_batch) */
      int r = srsb200_decode_tb_batch(e, &tb[first], m - first, 0);
      if (r != SRSB200_SUCCESS) {
        ret = r;
        for (uint32_t j = first; j < m; j++) {
          tb[j].ret = SRSB200_ERROR;
        }
      }
    }
  }
  for (uint32_t j = 0; j < m; j++) {
    uint32_t i = order[j];
    results[i] = tb[j].ret;
    if (cb_segm[i].tbs != 0 && cb_segm[i].C != 0 && tb[j].ret != SRSRAN_ERROR_INVALID_INPUTS && tb[j].ret != SRSB200_ERROR_NO_DEVICE) {
      softbuffer[i]->tb_crc = tb_crc[i] != 0;
      q[i]->avg_iterations  = tb[j].avg_iterations;
    }
  }
  free(tb);
  free(tb_crc);
  free(order);
  return ret == SRSB200_SUCCESS ? SRSRAN_SUCCESS : SRSRAN_ERROR;
}

/* ------------------------------------------------------------------ eNB uplink: srsran_ulsch_decode's data path */
/*
 * ulsch_deinterleave (sch.c:994-1021) + decode_tb (sch.c:509-573) of srsran_ulsch_decode (sch.c:1122-1193) in one device
 * submission: the interleaved LLRs q_bits go to the device once, are de-interleaved there (RI positions skipped) and the
 * transport block is decoded from g + e_offset. If g_bits != NULL its first nof_g_out values are returned (the CQI
 * decoder reads the front of the de-interleaved stream). ri_bits / nof_ri_bits are q->ack_ri_bits / Q'_ri * Qm.
 */
int srsran_b200_ulsch_decode_tb(srsran_sch_t*           q,
                                srsran_softbuffer_rx_t* softbuffer,
                                srsran_cbsegm_t*        cb_segm,
                                uint32_t                Qm,
                                uint32_t                rv,
                                int16_t*                q_bits,
                                uint32_t                H_prime_total,
                                uint32_t                N_pusch_symbs,
                                srsran_uci_bit_t*       ri_bits,
                                uint32_t                nof_ri_bits,
                                uint32_t                e_offset,
                                uint32_t                nof_e_bits,
                                int16_t*                g_bits,
                                uint32_t                nof_g_out,
                                uint8_t*                data)
{
  if (q == NULL || data == NULL || softbuffer == NULL || q_bits == NULL || cb_segm == NULL || Qm == 0 || (nof_ri_bits && ri_bits == NULL)) {
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint32_t* ri = nof_ri_bits ? malloc(nof_ri_bits * sizeof(uint32_t)) : NULL;
  for (uint32_t i = 0; i < nof_ri_bits; i++) {
    ri[i] = ri_bits[i].position;
  }
  uint8_t      tb_crc = 0;
  srsb200_tb_t tb;
  memset(&tb, 0, sizeof(tb));
  tb.tbs           = cb_segm->tbs;
  tb.Qm            = Qm;
  tb.rv            = rv;
  tb.nof_e_bits    = nof_e_bits;
  tb.buffer_f      = softbuffer->buffer_f;
  tb.sb_data       = softbuffer->data;
  tb.cb_crc        = (uint8_t*)softbuffer->cb_crc;
  tb.tb_crc        = &tb_crc;
  tb.max_cb        = softbuffer->max_cb;
  tb.data          = data;
  tb.q_bits        = q_bits;
  tb.H_prime_total = H_prime_total;
  tb.N_pusch_symbs = N_pusch_symbs;
  tb.ri_positions  = ri;
  tb.nof_ri_bits   = nof_ri_bits;
  tb.e_offset      = e_offset;
  tb.g_bits        = g_bits;
  tb.nof_g_out     = g_bits ? nof_g_out : 0;
  int ret = srsb200_decode_tb(engine_for(softbuffer), &tb, q->max_iterations);
  free(ri);
  if (ret == SRSB200_ERROR_NO_DEVICE) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    return SRSRAN_ERROR;
  }
  if (cb_segm->tbs != 0 && cb_segm->C != 0 && ret != SRSRAN_ERROR_INVALID_INPUTS) {
    softbuffer->tb_crc = tb_crc != 0;
    q->avg_iterations  = tb.avg_iterations;
  }
  return ret;
}

/* ulsch_deinterleave alone (for grants that carry CQI: the CQI decoder needs g_bits on the host before decode_tb) */
int srsran_b200_ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits,
                                   srsran_uci_bit_t* ri_bits, uint32_t nof_ri_bits)
{
  uint32_t* ri = nof_ri_bits ? malloc(nof_ri_bits * sizeof(uint32_t)) : NULL;
  for (uint32_t i = 0; i < nof_ri_bits; i++) {
    ri[i] = ri_bits[i].position;
  }
  int ret = srsb200_ulsch_deinterleave(engine(), q_bits, Qm, H_prime_total, N_pusch_symbs, g_bits, ri, nof_ri_bits);
  free(ri);
  return ret == SRSB200_SUCCESS ? SRSRAN_SUCCESS : (ret == SRSB200_ERROR_INVALID_INPUTS ? SRSRAN_ERROR_INVALID_INPUTS : SRSRAN_ERROR);
}

/* ------------------------------------------------------------------ transmit side: sch.c encode_tb_off */
/* dst bits [off, off+n) <- src bits [0, n), MSB first; every other bit of dst is preserved (what srsran_bit_copy does) */
static void splice_bits(uint8_t* dst, uint32_t off, const uint8_t* src, uint32_t n)
{
  if ((off & 7u) == 0) {
    memcpy(&dst[off / 8], src, n / 8);
    if (n & 7u) {
      uint8_t mask = (uint8_t)(0xff00u >> (n & 7u));
      dst[off / 8 + n / 8] = (uint8_t)((dst[off / 8 + n / 8] & ~mask) | (src[n / 8] & mask));
    }
    return;
  }
  for (uint32_t i = 0; i < n; i++) {
    uint32_t p   = off + i;
    uint8_t  bit = (src[i / 8] >> (7 - (i & 7u))) & 1u;
    dst[p / 8]   = (uint8_t)((dst[p / 8] & ~(0x80u >> (p & 7u))) | (bit ? (0x80u >> (p & 7u)) : 0));
  }
}

/*
 * data == NULL (sch.c:305 `if (data)`): the reference then re-reads the circular buffers softbuffer->buffer_b that an earlier
 * call WITH data filled, i.e. it retransmits the payload of that call. The device encoder works from the payload, so the shim
 * keeps the payload of the last call with data in the (otherwise unused) buffer_b arrays of the same soft buffer - same
 * lifetime and ownership as the reference's state, no allocation: an 8-byte header in buffer_b[0], then PAYLOAD_CHUNK bytes
 * per code-block buffer.
 */
#define PAYLOAD_MAGIC 0xB2005EEDu
#define PAYLOAD_CHUNK 16384u
static int payload_fits(const srsran_softbuffer_tx_t* sb, uint32_t nbytes)
{
  if (sb->buffer_b == NULL || sb->max_cb_size < PAYLOAD_CHUNK + 8) {
    return 0;
  }
  return (nbytes + PAYLOAD_CHUNK - 1) / PAYLOAD_CHUNK <= sb->max_cb;
}
static void payload_store(srsran_softbuffer_tx_t* sb, uint32_t tbs, const uint8_t* data)
{
  uint32_t nbytes = tbs / 8;
  if (!payload_fits(sb, nbytes) || sb->buffer_b[0] == NULL) {
    return;
  }
  for (uint32_t i = 0, off = 0; off < nbytes; i++, off += PAYLOAD_CHUNK) {
    if (sb->buffer_b[i] == NULL) {
      return;
    }
    uint32_t n = nbytes - off < PAYLOAD_CHUNK ? nbytes - off : PAYLOAD_CHUNK;
    memcpy(sb->buffer_b[i] + 8, &data[off], n);
  }
  uint32_t hdr[2] = {PAYLOAD_MAGIC, tbs};
  memcpy(sb->buffer_b[0], hdr, sizeof(hdr));
}
/* -> malloc'd copy of the payload stored for this transport block size, or NULL */
static uint8_t* payload_load(const srsran_softbuffer_tx_t* sb, uint32_t tbs)
{
  uint32_t nbytes = tbs / 8, hdr[2];
  if (!payload_fits(sb, nbytes) || sb->buffer_b[0] == NULL) {
    return NULL;
  }
  memcpy(hdr, sb->buffer_b[0], sizeof(hdr));
  if (hdr[0] != PAYLOAD_MAGIC || hdr[1] != tbs) {
    return NULL;
  }
  uint8_t* p = malloc(nbytes + 8);
  for (uint32_t i = 0, off = 0; p && off < nbytes; i++, off += PAYLOAD_CHUNK) {
    uint32_t n = nbytes - off < PAYLOAD_CHUNK ? nbytes - off : PAYLOAD_CHUNK;
    memcpy(&p[off], sb->buffer_b[i] + 8, n);
  }
  return p;
}

/*
 * Drop-in body of sch.c's static encode_tb_off(): same arguments and return codes. CRC attach, turbo encoding and rate
 * matching of all code blocks run on the device as one submission; the bits land at bit offset w_offset of e_bits and
 * the rest of e_bits is preserved, as srsran_rm_turbo_tx_lut + srsran_bit_copy leave it. The transmit soft buffer
 * (softbuffer->buffer_b, the per-block circular buffers) holds the payload instead of the coded circular buffers (above):
 * every redundancy version is produced from the payload, with or without `data`.
 */
int srsran_b200_encode_tb(srsran_sch_t*           q,
                          srsran_softbuffer_tx_t* softbuffer,
                          srsran_cbsegm_t*        cb_segm,
                          uint32_t                Qm,
                          uint32_t                rv,
                          uint32_t                nof_e_bits,
                          uint8_t*                data,
                          uint8_t*                e_bits,
                          uint32_t                w_offset)
{
  if (q == NULL || e_bits == NULL || cb_segm == NULL || softbuffer == NULL) {
    ERROR("Invalid parameters: e_bits=%d, cb_segm=%d, softbuffer=%d", e_bits != 0, cb_segm != 0, softbuffer != 0);
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  uint8_t* kept = NULL;
  if (data == NULL && cb_segm->C > 0 && !cb_segm->F && Qm) {
    kept = payload_load(softbuffer, cb_segm->tbs); /* HARQ retransmission: the payload of the last call with data */
    if (kept == NULL) {
      ERROR("srsran_b200: retransmission without payload, and this soft buffer holds none for tbs=%d", cb_segm->tbs);
      return SRSRAN_ERROR;
    }
    data = kept;
  } else if (data != NULL && cb_segm->C > 0 && !cb_segm->F) {
    payload_store(softbuffer, cb_segm->tbs, data);
  }
  uint32_t nbytes = (nof_e_bits + 7) / 8;
  uint8_t* tmp    = w_offset || (Qm && nof_e_bits % Qm) || (nof_e_bits & 7u) ? malloc(nbytes + 8) : e_bits;
  if (!tmp) {
    free(kept);
    return SRSRAN_ERROR;
  }
  srsb200_tb_tx_t tb;
  memset(&tb, 0, sizeof(tb));
  tb.tbs        = cb_segm->tbs;
  tb.Qm         = Qm;
  tb.rv         = rv;
  tb.nof_e_bits = nof_e_bits;
  tb.max_cb     = softbuffer->max_cb;
  tb.data       = data ? data : (const uint8_t*)"";
  tb.e_bits     = tmp;
  int ret = srsb200_encode_tb(engine_for(softbuffer), &tb);
  if (ret == SRSB200_ERROR_NO_DEVICE) {
    ERROR("srsran_b200: %s", srsb200_last_error());
    ret = SRSRAN_ERROR;
  }
  if (tmp != e_bits) {
    if (ret == SRSRAN_SUCCESS && cb_segm->C > 0 && Qm) {
      splice_bits(e_bits, w_offset, tmp, Qm * (nof_e_bits / Qm)); /* sum of the per-block E, sch.c:299-303 */
    }
    free(tmp);
  }
  free(kept);
  return ret;
}

/* ------------------------------------------------------------------ self-test hook (tests/test_shim.py, ctypes) */
/*
 * Runs srsran_b200_decode_tb through REAL reference structs (srsran_sch_t, srsran_softbuffer_rx_t, srsran_cbsegm_t built
 * by the reference's own srsran_cbsegm is not linked here, so the caller passes C/K1/... as computed by the library).
 * Soft-buffer arrays are passed flat: buffer_f[max_cb][SOFTBUFFER_SIZE], sb_data[max_cb][SOFTBUFFER_SIZE/8].
 */
int srsran_b200_selftest_decode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, int16_t* e_bits, uint32_t max_iterations,
                                   uint32_t max_cb, int16_t* buffer_f, uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc, uint8_t* data,
                                   float* avg_iterations)
{
  srsran_sch_t* q = calloc(1, sizeof(srsran_sch_t));
  if (!q) {
    return SRSRAN_ERROR;
  }
  q->max_iterations = max_iterations ? max_iterations : 10;
  srsran_softbuffer_rx_t sb;
  memset(&sb, 0, sizeof(sb));
  sb.max_cb      = max_cb;
  sb.max_cb_size = SOFTBUFFER_SIZE;
  sb.buffer_f    = calloc(max_cb, sizeof(int16_t*));
  sb.data        = calloc(max_cb, sizeof(uint8_t*));
  sb.cb_crc      = calloc(max_cb, sizeof(bool));
  for (uint32_t i = 0; i < max_cb; i++) {
    sb.buffer_f[i] = &buffer_f[(size_t)i * SOFTBUFFER_SIZE];
    sb.data[i]     = &sb_data[(size_t)i * (SOFTBUFFER_SIZE / 8)];
    sb.cb_crc[i]   = cb_crc[i] != 0;
  }
  uint32_t        sg[8];
  srsran_cbsegm_t seg;
  memset(&seg, 0, sizeof(seg));
  srsb200_cbsegm(tbs, sg);
  seg.F = sg[0]; seg.C = sg[1]; seg.K1 = sg[2]; seg.K2 = sg[3]; seg.K1_idx = sg[4]; seg.K2_idx = sg[5]; seg.C1 = sg[6]; seg.C2 = sg[7];
  seg.tbs = tbs; seg.L_tb = 24; seg.L_cb = 24;
  int ret = srsran_b200_decode_tb(q, &sb, &seg, Qm, rv, nof_e_bits, e_bits, data);
  for (uint32_t i = 0; i < max_cb; i++) {
    cb_crc[i] = sb.cb_crc[i] ? 1 : 0;
  }
  *tb_crc         = sb.tb_crc ? 1 : 0;
  *avg_iterations = q->avg_iterations;
  free(sb.buffer_f);
  free(sb.data);
  free(sb.cb_crc);
  free(q);
  return ret;
}

/* srsran_b200_dlsch_decode2_symbols through a real srsran_pdsch_cfg_t: what the patched pdsch.c calls (one codeword, tb_idx 0) */
int srsran_b200_selftest_dlsch_symbols(uint32_t tbs, int mod, uint32_t rv, uint32_t nof_re, uint32_t nof_tb, uint32_t nof_layers, float* symbols,
                                       uint16_t rnti, int codeword_idx, uint32_t nslot, uint32_t cell_id, uint32_t max_iterations, uint32_t max_cb,
                                       int16_t* buffer_f, uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc, uint8_t* data, float* avg_iterations)
{
  static const uint32_t bps[5] = {1, 2, 4, 6, 8};
  srsran_sch_t* q = calloc(1, sizeof(srsran_sch_t));
  if (!q || mod < 0 || mod > 4) {
    free(q);
    return SRSRAN_ERROR;
  }
  q->max_iterations = max_iterations ? max_iterations : 10;
  srsran_softbuffer_rx_t sb;
  memset(&sb, 0, sizeof(sb));
  sb.max_cb      = max_cb;
  sb.max_cb_size = SOFTBUFFER_SIZE;
  sb.buffer_f    = calloc(max_cb, sizeof(int16_t*));
  sb.data        = calloc(max_cb, sizeof(uint8_t*));
  sb.cb_crc      = calloc(max_cb, sizeof(bool));
  for (uint32_t i = 0; i < max_cb; i++) {
    sb.buffer_f[i] = &buffer_f[(size_t)i * SOFTBUFFER_SIZE];
    sb.data[i]     = &sb_data[(size_t)i * (SOFTBUFFER_SIZE / 8)];
    sb.cb_crc[i]   = cb_crc[i] != 0;
  }
  srsran_pdsch_cfg_t* cfg = calloc(1, sizeof(srsran_pdsch_cfg_t));
  cfg->grant.nof_tb         = nof_tb;
  cfg->grant.nof_layers     = nof_layers;
  cfg->grant.nof_re         = nof_re;
  cfg->grant.tb[0].tbs      = (int)tbs;
  cfg->grant.tb[0].mod      = (srsran_mod_t)mod;
  cfg->grant.tb[0].rv       = (int)rv;
  cfg->grant.tb[0].nof_bits = nof_re * bps[mod];
  cfg->grant.tb[0].enabled  = true;
  cfg->softbuffers.rx[0]    = &sb;
  cfg->rnti                 = rnti;
  int ret = SRSRAN_ERROR_INVALID_INPUTS;
  if (srsran_b200_dlsch_takes_symbols(q, cfg, 0)) {
    ret = srsran_b200_dlsch_decode2_symbols(q, cfg, (cf_t*)symbols, srsran_b200_pdsch_c_init(rnti, codeword_idx, nslot, cell_id), data, 0, nof_layers);
  }
  for (uint32_t i = 0; i < max_cb; i++) {
    cb_crc[i] = sb.cb_crc[i] ? 1 : 0;
  }
  *tb_crc         = sb.tb_crc ? 1 : 0;
  *avg_iterations = q->avg_iterations;
  free(sb.buffer_f);
  free(sb.data);
  free(sb.cb_crc);
  free(cfg);
  free(q);
  return ret;
}

size_t srsran_b200_selftest_sizeof_tdec(void)
{
  return sizeof(srsran_tdec_t);
}

/* srsran_b200_encode_tb through real srsran_sch_t / srsran_softbuffer_tx_t / srsran_cbsegm_t objects */
int srsran_b200_selftest_encode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, uint8_t* data, uint8_t* e_bits, uint32_t w_offset,
                                   uint32_t max_cb)
{
  srsran_sch_t* q = calloc(1, sizeof(srsran_sch_t));
  if (!q) {
    return SRSRAN_ERROR;
  }
  srsran_softbuffer_tx_t sb;
  memset(&sb, 0, sizeof(sb));
  sb.max_cb = max_cb;
  uint32_t        sg[8];
  srsran_cbsegm_t seg;
  memset(&seg, 0, sizeof(seg));
  srsb200_cbsegm(tbs, sg);
  seg.F = sg[0]; seg.C = sg[1]; seg.K1 = sg[2]; seg.K2 = sg[3]; seg.K1_idx = sg[4]; seg.K2_idx = sg[5]; seg.C1 = sg[6]; seg.C2 = sg[7];
  seg.tbs = tbs; seg.L_tb = 24; seg.L_cb = 24;
  int ret = srsran_b200_encode_tb(q, &sb, &seg, Qm, rv, nof_e_bits, data, e_bits, w_offset);
  free(q);
  return ret;
}

/* srsran_b200_ulsch_decode_tb through real reference structs; RI positions are given as plain indices */
int srsran_b200_selftest_ulsch_decode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, int16_t* q_bits, uint32_t H_prime_total, uint32_t N_pusch_symbs,
                                         const uint32_t* ri_positions, uint32_t nof_ri_bits, uint32_t e_offset, uint32_t nof_e_bits, int16_t* g_bits,
                                         uint32_t nof_g_out, uint32_t max_iterations, uint32_t max_cb, int16_t* buffer_f, uint8_t* sb_data,
                                         uint8_t* cb_crc, uint8_t* tb_crc, uint8_t* data)
{
  srsran_sch_t* q = calloc(1, sizeof(srsran_sch_t));
  if (!q) {
    return SRSRAN_ERROR;
  }
  q->max_iterations = max_iterations ? max_iterations : 10;
  srsran_softbuffer_rx_t sb;
  memset(&sb, 0, sizeof(sb));
  sb.max_cb      = max_cb;
  sb.max_cb_size = SOFTBUFFER_SIZE;
  sb.buffer_f    = calloc(max_cb, sizeof(int16_t*));
  sb.data        = calloc(max_cb, sizeof(uint8_t*));
  sb.cb_crc      = calloc(max_cb, sizeof(bool));
  for (uint32_t i = 0; i < max_cb; i++) {
    sb.buffer_f[i] = &buffer_f[(size_t)i * SOFTBUFFER_SIZE];
    sb.data[i]     = &sb_data[(size_t)i * (SOFTBUFFER_SIZE / 8)];
    sb.cb_crc[i]   = cb_crc[i] != 0;
  }
  srsran_uci_bit_t* ri = calloc(nof_ri_bits + 1, sizeof(srsran_uci_bit_t));
  for (uint32_t i = 0; i < nof_ri_bits; i++) {
    ri[i].position = ri_positions[i];
  }
  uint32_t        sg[8];
  srsran_cbsegm_t seg;
  memset(&seg, 0, sizeof(seg));
  srsb200_cbsegm(tbs, sg);
  seg.F = sg[0]; seg.C = sg[1]; seg.K1 = sg[2]; seg.K2 = sg[3]; seg.K1_idx = sg[4]; seg.K2_idx = sg[5]; seg.C1 = sg[6]; seg.C2 = sg[7];
  seg.tbs = tbs; seg.L_tb = 24; seg.L_cb = 24;
  int ret = srsran_b200_ulsch_decode_tb(q, &sb, &seg, Qm, rv, q_bits, H_prime_total, N_pusch_symbs, ri, nof_ri_bits, e_offset, nof_e_bits, g_bits, nof_g_out,
                                        data);
  for (uint32_t i = 0; i < max_cb; i++) {
    cb_crc[i] = sb.cb_crc[i] ? 1 : 0;
  }
  *tb_crc = sb.tb_crc ? 1 : 0;
  free(ri);
  free(sb.buffer_f);
  free(sb.data);
  free(sb.cb_crc);
  free(q);
  return ret;
}
