/*
 * ref_shim.c - TEST INFRASTRUCTURE ONLY.
 *
 * Flat, ctypes-friendly entry points around the UNMODIFIED srsRAN 4G reference primitives, compiled by oracle/Makefile
 * from the sources where they lie under /root/reference (nothing from the reference is copied into this repo; this
 * file is our own glue and only *calls* reference functions). The resulting oracle/_ref/libsrsran_ref.so is used
 *   - by tests/ to pin the clean-room restatement (oracle/turbo_oracle.c) and to generate tests/golden/ fixtures,
 *   - by bench.py --impl reference / the cpu_baseline leg as the timed CPU baseline (AVX2 windowed decoder).
 * It is never loaded by the product library (srsran_4g_b200/).
 *
 * Oracle definition: srsran_tdec_init_manual(.., SRSRAN_TDEC_GENERIC) + natural input layout
 * (srsran_rm_turbo_rx_lut_(.., false)) + the decode_tb_cb loop semantics (lib/src/phy/phch/sch.c:371-494) +
 * srsran_crc_checksum_byte.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "srsran/phy/fec/cbsegm.h"
#include "srsran/phy/fec/crc.h"
#include "srsran/phy/fec/softbuffer.h"
#include "srsran/phy/fec/turbo/rm_turbo.h"
#include "srsran/phy/fec/turbo/tc_interl.h"
#include "srsran/phy/fec/turbo/turbocoder.h"
#include "srsran/phy/fec/turbo/turbodecoder.h"
#include "srsran/phy/fec/turbo/turbodecoder_gen.h"

#define CRC24A 0x1864CFB
#define CRC24B 0x1800063
#define MAX_K 6144
#define MAX_L (3 * MAX_K + 12)

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static int             g_init = 0;
static srsran_tcod_t   g_tcod;
static srsran_tdec_t   g_tdec_gen;

int ref_init(void)
{
  pthread_mutex_lock(&g_lock);
  if (!g_init) {
    srsran_rm_turbo_gentables();
    if (srsran_tcod_init(&g_tcod, MAX_K)) {
      pthread_mutex_unlock(&g_lock);
      return -1;
    }
    if (srsran_tdec_init_manual(&g_tdec_gen, MAX_K, SRSRAN_TDEC_GENERIC)) {
      pthread_mutex_unlock(&g_lock);
      return -1;
    }
    g_init = 1;
  }
  pthread_mutex_unlock(&g_lock);
  return 0;
}

int ref_cbsize(uint32_t idx) { return srsran_cbsegm_cbsize(idx); }
int ref_cbindex(uint32_t K) { return srsran_cbsegm_cbindex(K); }

/* out[12] = F C K1 K2 K1_idx K2_idx C1 C2 tbs L_tb L_cb Z */
int ref_cbsegm(uint32_t tbs, uint32_t* out)
{
  srsran_cbsegm_t s;
  memset(&s, 0, sizeof(s));
  int ret = srsran_cbsegm(&s, tbs);
  out[0]  = s.F;
  out[1]  = s.C;
  out[2]  = s.K1;
  out[3]  = s.K2;
  out[4]  = s.K1_idx;
  out[5]  = s.K2_idx;
  out[6]  = s.C1;
  out[7]  = s.C2;
  out[8]  = s.tbs;
  out[9]  = s.L_tb;
  out[10] = s.L_cb;
  out[11] = s.Z;
  return ret;
}

int ref_qpp(uint32_t K, uint16_t* fwd, uint16_t* rev)
{
  srsran_tc_interl_t t;
  if (srsran_tc_interl_init(&t, K)) {
    return -1;
  }
  int ret = srsran_tc_interl_LTE_gen(&t, K);
  if (!ret) {
    memcpy(fwd, t.forward, sizeof(uint16_t) * K);
    memcpy(rev, t.reverse, sizeof(uint16_t) * K);
  }
  srsran_tc_interl_free(&t);
  return ret;
}

uint32_t ref_crc_bytes(uint32_t poly, int order, const uint8_t* data, int nbits)
{
  srsran_crc_t c;
  srsran_crc_init(&c, poly, order);
  return srsran_crc_checksum_byte(&c, data, nbits);
}

/* CRC over unpacked bits (one bit per byte), any length: srsran_crc_checksum (crc.c:94-144) */
uint32_t ref_crc_bits(uint32_t poly, int order, uint8_t* bits, int nbits)
{
  srsran_crc_t c;
  srsran_crc_init(&c, poly, order);
  return srsran_crc_checksum(&c, bits, nbits);
}

/* natural-layout rate de-matching: output[T[i mod L]] += input[i] */
int ref_rm_rx(int16_t* e, int16_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  ref_init();
  return srsran_rm_turbo_rx_lut_(e, buf, E, cb_idx, rv, false);
}

/* the reference's production (AUTO) layout */
int ref_rm_rx_auto(int16_t* e, int16_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  ref_init();
  return srsran_rm_turbo_rx_lut(e, buf, E, cb_idx, rv);
}

/* Recover the natural de-rate-matching table T[0..L) by impulse response of the reference function. */
int ref_rm_table(uint32_t cb_idx, uint32_t rv, uint16_t* table)
{
  ref_init();
  int K = srsran_cbsegm_cbsize(cb_idx);
  if (K < 0 || rv >= 4) {
    return -1;
  }
  uint32_t L   = 3 * K + 12;
  int16_t* e   = calloc(L + 64, sizeof(int16_t));
  int16_t* buf = calloc(L + 256, sizeof(int16_t));
  for (uint32_t i = 0; i < L; i++) {
    e[i] = (int16_t)(i + 1);
  }
  int ret = srsran_rm_turbo_rx_lut_(e, buf, L, cb_idx, rv, false);
  for (uint32_t i = 0; i < L; i++) {
    table[i] = 0xffff;
  }
  for (uint32_t j = 0; j < L; j++) {
    if (buf[j] <= 0) {
      ret = -2; /* not a permutation */
    } else {
      table[buf[j] - 1] = (uint16_t)j;
    }
  }
  free(e);
  free(buf);
  return ret;
}

/* bitwise encoder: in[K] bits -> out[3K+12] bits (natural order) */
int ref_tcod_encode(uint8_t* in, uint8_t* out, uint32_t K)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int r = srsran_tcod_encode(&g_tcod, in, out, K);
  pthread_mutex_unlock(&g_lock);
  return r;
}

/* bitwise rate matching (tx): coded[3K+12] bits -> e[E] bits */
int ref_rm_tx(uint8_t* coded, uint32_t K, uint8_t* e, uint32_t E, uint32_t rv)
{
  /* the reference fills its circular buffer only when rv == 0 (rm_turbo.c:1016) and reuses it for later rv */
  uint8_t* w = calloc(3 * (MAX_K + 64) + 64, 1);
  int      r = srsran_rm_turbo_tx(w, 3 * (MAX_K + 64), coded, 3 * K + 12, e, E, 0);
  if (!r && rv != 0) {
    r = srsran_rm_turbo_tx(w, 3 * (MAX_K + 64), coded, 3 * K + 12, e, E, rv);
  }
  free(w);
  return r;
}

static srsran_tdec_impl_type_t impl_of(int impl)
{
  switch (impl) {
    case 1:
      return SRSRAN_TDEC_GENERIC;
    case 3:
      return SRSRAN_TDEC_SSE_WINDOW;
    case 5:
      return SRSRAN_TDEC_AVX_WINDOW;
    default:
      return SRSRAN_TDEC_AUTO;
  }
}

/*
 * Run nof_iter half-iterations of one code block (natural input layout), recording the hard decision after every
 * half-iteration (out_bytes[it][K/8]) and optionally the soft arrays of the generic decoder after every half-iteration
 * (dump[it][3][K] = ext1, ext2, app1).
 */
int ref_tdec_trace(int impl, uint32_t K, int16_t* in, uint32_t nof_iter, uint8_t* out_bytes, int16_t* dump)
{
  srsran_tdec_t h;
  if (srsran_tdec_init_manual(&h, K, impl_of(impl))) {
    return -1;
  }
  srsran_tdec_force_not_sb(&h);
  if (srsran_tdec_new_cb(&h, K)) {
    srsran_tdec_free(&h);
    return -2;
  }
  for (uint32_t it = 0; it < nof_iter; it++) {
    srsran_tdec_iteration(&h, in, &out_bytes[it * (K / 8)]);
    if (dump) {
      memcpy(&dump[(it * 3 + 0) * K], h.ext1, sizeof(int16_t) * K);
      memcpy(&dump[(it * 3 + 1) * K], h.ext2, sizeof(int16_t) * K);
      memcpy(&dump[(it * 3 + 2) * K], h.app1, sizeof(int16_t) * K);
    }
  }
  srsran_tdec_free(&h);
  return 0;
}

/*
 * The 8-BIT mode (turbodecoder.c:458-484,551-555): srsran_tdec_iteration_8bit in AUTO mode on natural-order int8 input
 * (srsran_tdec_force_not_sb => tdec_win*_extract_input does the window interleaving). dump[it][3][K] = ext1, ext2, app1 of the
 * decoder object after every half-iteration, brought back from the window-interleaved storage to natural order
 * (element n lives at (n % S) * NW + n / S, S = K / NW: turbodecoder_win.h:883-921). Returns the number of windows the
 * reference used (0 = it left 8-bit arithmetic for this K: turbodecoder.c:443-476) or < 0.
 */
int ref_tdec8_trace(uint32_t K, int8_t* in, uint32_t nof_iter, uint8_t* out_bytes, int8_t* dump)
{
  srsran_tdec_t h;
  if (srsran_tdec_init(&h, K)) {
    return -1;
  }
  srsran_tdec_force_not_sb(&h);
  if (srsran_tdec_new_cb(&h, K)) {
    srsran_tdec_free(&h);
    return -2;
  }
  uint32_t NW = srsran_tdec_autoimp_get_subblocks_8bit(K);
  uint32_t nw = (NW == 32 || NW == 16) ? NW : 0;
  for (uint32_t it = 0; it < nof_iter; it++) {
    srsran_tdec_iteration_8bit(&h, in, &out_bytes[it * (K / 8)]);
    if (dump && nw) {
      const int8_t* src[3] = {(const int8_t*)h.ext1, (const int8_t*)h.ext2, (const int8_t*)h.app1};
      uint32_t      S      = K / nw;
      for (int a = 0; a < 3; a++) {
        for (uint32_t n = 0; n < K; n++) {
          dump[(it * 3 + a) * K + n] = src[a][(n % S) * nw + n / S];
        }
      }
    }
  }
  srsran_tdec_free(&h);
  return (int)nw;
}

/* srsran_rm_turbo_rx_lut_8bit itself: writes the sub-block layout of srsran_tdec_autoimp_get_subblocks_8bit(K) */
int ref_rm_rx8(int8_t* e, int8_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  ref_init();
  return srsran_rm_turbo_rx_lut_8bit(e, buf, E, cb_idx, rv);
}

/* srsran_tdec_run_all on one CB */
int ref_tdec_run_all(int impl, uint32_t K, int16_t* in, uint32_t nof_iter, uint8_t* out_bytes)
{
  srsran_tdec_t h;
  if (srsran_tdec_init_manual(&h, K, impl_of(impl))) {
    return -1;
  }
  srsran_tdec_force_not_sb(&h);
  int r = srsran_tdec_new_cb(&h, K);
  if (!r) {
    r = srsran_tdec_run_all(&h, in, out_bytes, nof_iter, K);
  }
  srsran_tdec_free(&h);
  return r;
}

/* one MAP decode through the reference's vtable seam (turbodecoder_impl.h:54-60), generic impl */
int ref_map_gen(uint32_t K, int16_t* input, int16_t* app, int16_t* parity, int16_t* output)
{
  void* hh = NULL;
  if (tdec_gen_init(&hh, K) < 0) {
    return -1;
  }
  tdec_gen_dec(hh, input, app, parity, output, K);
  tdec_gen_free(hh);
  return 0;
}

/*
 * The decode_tb_cb / decode_tb semantics (sch.c:371-494, 509-573) with the oracle substitutions:
 * GENERIC decoder + natural-layout rate de-matching. HARQ state is passed as flat arrays:
 *   buffer_f[C][18600] int16, sb_data[C][18600/8] u8, cb_crc[C] u8 (in/out), tb_crc (out)
 * Returns the decode_tb return code (0 ok, -1 CRC fail, -2 invalid). cb_noi[C] receives per-CB half-iterations run in
 * this call (0 for CBs skipped because cb_crc was already set); *avg_iterations mirrors q->avg_iterations.
 */
int ref_decode_tb(uint32_t tbs,
                  uint32_t Qm,
                  uint32_t rv,
                  uint32_t nof_e_bits,
                  int16_t* e_bits,
                  uint32_t max_iterations,
                  int16_t* buffer_f,
                  uint8_t* sb_data,
                  uint8_t* cb_crc,
                  uint8_t* tb_crc,
                  uint8_t* data,
                  uint32_t* cb_noi,
                  float*   avg_iterations)
{
  ref_init();
  srsran_cbsegm_t seg;
  if (srsran_cbsegm(&seg, tbs)) {
    return -1;
  }
  if (Qm == 0) {
    return -2;
  }
  if (seg.tbs == 0 || seg.C == 0) {
    return 0;
  }
  if (seg.F) {
    return -2;
  }
  srsran_crc_t crc_tb, crc_cb;
  srsran_crc_init(&crc_tb, CRC24A, 24);
  srsran_crc_init(&crc_cb, CRC24B, 24);
  if (max_iterations == 0) {
    max_iterations = 10;
  }

  srsran_tdec_t* dec = &g_tdec_gen;
  pthread_mutex_lock(&g_lock);
  float avg = 0;
  for (uint32_t cb_idx = 0; cb_idx < seg.C; cb_idx++) {
    uint32_t cb_len = cb_idx < seg.C1 ? seg.K1 : seg.K2;
    uint32_t rlen   = seg.C == 1 ? cb_len : (cb_len - 24);
    cb_noi[cb_idx]  = 0;
    if (!cb_crc[cb_idx]) {
      uint32_t cb_len_idx = cb_idx < seg.C1 ? seg.K1_idx : seg.K2_idx;
      uint32_t Gp         = nof_e_bits / Qm;
      uint32_t gamma      = Gp % seg.C;
      uint32_t n_e        = Qm * (Gp / seg.C);
      uint32_t rp         = cb_idx * n_e;
      uint32_t n_e2       = n_e;
      if (cb_idx > seg.C - gamma) {
        n_e2 = n_e + Qm;
        rp   = (seg.C - gamma) * n_e + (cb_idx - (seg.C - gamma)) * n_e2;
      }
      int16_t* buf = &buffer_f[(size_t)cb_idx * SOFTBUFFER_SIZE];
      srsran_rm_turbo_rx_lut_(&e_bits[rp], buf, n_e2, cb_len_idx, rv, false);
      srsran_tdec_new_cb(dec, cb_len);
      bool     early_stop = false;
      uint32_t noi        = 0;
      do {
        srsran_tdec_iteration(dec, buf, &data[cb_idx * rlen / 8]);
        avg++;
        noi++;
        uint32_t      len_crc = seg.C > 1 ? cb_len : seg.tbs + 24;
        srsran_crc_t* crc_ptr = seg.C > 1 ? &crc_cb : &crc_tb;
        if (!srsran_crc_checksum_byte(crc_ptr, &data[cb_idx * rlen / 8], len_crc) && noi >= 2) {
          cb_crc[cb_idx] = 1;
          early_stop     = true;
        }
      } while (noi < max_iterations && !early_stop);
      cb_noi[cb_idx] = noi;
    } else {
      memcpy(&data[cb_idx * rlen / 8], &sb_data[(size_t)cb_idx * (SOFTBUFFER_SIZE / 8)], rlen / 8);
    }
  }
  pthread_mutex_unlock(&g_lock);

  bool tb_ok = true;
  for (uint32_t i = 0; i < seg.C && tb_ok; i++) {
    tb_ok = cb_crc[i];
  }
  *tb_crc = tb_ok;
  if (!tb_ok) {
    for (uint32_t i = 0; i < seg.C; i++) {
      if (cb_crc[i]) {
        uint32_t cb_len = i < seg.C1 ? seg.K1 : seg.K2;
        uint32_t rlen   = seg.C == 1 ? cb_len : (cb_len - 24);
        memcpy(&sb_data[(size_t)i * (SOFTBUFFER_SIZE / 8)], &data[i * rlen / 8], rlen / 8);
      }
    }
  }
  *avg_iterations = avg / (float)seg.C;
  if (!tb_ok) {
    return -1;
  }
  if (seg.C == 1) {
    return 0;
  }
  if (srsran_crc_match_byte(&crc_tb, data, seg.tbs)) {
    return 0;
  }
  for (uint32_t i = 0; i < seg.C; i++) {
    cb_crc[i] = 0; /* srsran_softbuffer_rx_reset_cb_crc */
  }
  return -1;
}

/* ---------------- timed CPU baseline: batch of equal-K code blocks over pthreads ---------------- */
typedef struct {
  int       impl;
  uint32_t  K;
  int16_t*  in; /* [n][3K+12] natural layout */
  uint8_t*  out; /* [n][K/8] */
  uint8_t*  noi; /* [n] */
  uint8_t*  crc_ok; /* [n] */
  uint32_t  first, last;
  uint32_t  max_iter;
  int       early_stop; /* 0: run exactly max_iter; 1: CRC24B early stop, min 2 */
  int       core;
  double    t_start, t_end;
  int       err;
  pthread_barrier_t* barrier;
  int       sb_layout; /* 1: decode from the sub-block soft-buffer layout prepared outside the timed region (production path) */
} job_t;

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define SB_STRIDE 18624 /* SOFTBUFFER_SIZE rounded up to a multiple of 32 int16 (64 bytes) */
static void* job_run(void* arg)
{
  job_t* j = (job_t*)arg;
  if (j->core >= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(j->core, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  srsran_tdec_t h;
  srsran_crc_t  crc;
  srsran_crc_init(&crc, CRC24B, 24);
  /* decoder construction (interleaver tables for all sizes) is set-up cost, kept outside the timed region */
  /* (the sub-block input layout is only read in AUTO mode, where current_dec > 0: turbodecoder.c:486-493, turbodecoder_iter.h:90;
   *  for K > 800 and K % 16 == 0 AUTO selects the same AVX2 windowed decoder as impl 5) */
  if (srsran_tdec_init_manual(&h, j->K, j->sb_layout ? SRSRAN_TDEC_AUTO : impl_of(j->impl))) {
    j->err = -1;
    pthread_barrier_wait(j->barrier);
    return NULL;
  }
  uint32_t L = 3 * j->K + 12;
  int16_t* sb = NULL;
  if (j->sb_layout) {
    /* production input (sch.c:415-433): the rate de-matcher writes the soft buffer in the decoder's sub-block layout
     * (srsran_tdec_autoimp_get_subblocks, rm_turbo.c:413) and srsran_tdec_iteration reads it as it is - no re-layout inside
     * the decoder. Built here, OUTSIDE the timed region: e-bits of a full rv-0 transmission that
     * de-rate-match to exactly the given natural-order LLRs (e[i] = d[T[i]], T the reference's own table). */
    int       cb_idx = srsran_cbsegm_cbindex(j->K);
    uint16_t* T      = calloc(L, sizeof(uint16_t));
    int16_t*  e      = calloc(L + 64, sizeof(int16_t));
    /* (the SIMD decoder does aligned loads from the soft buffer: the reference allocates it with srsran_vec_i16_malloc) */
    if (posix_memalign((void**)&sb, 64, ((size_t)(j->last - j->first) * SB_STRIDE + 64) * sizeof(int16_t))) {
      sb = NULL;
    } else {
      memset(sb, 0, ((size_t)(j->last - j->first) * SB_STRIDE + 64) * sizeof(int16_t));
    }
    if (!T || !e || !sb || cb_idx < 0 || ref_rm_table((uint32_t)cb_idx, 0, T)) {
      j->err = -1;
    } else {
      for (uint32_t n = j->first; n < j->last; n++) {
        const int16_t* d = &j->in[(size_t)n * L];
        for (uint32_t i = 0; i < L; i++) {
          e[i] = d[T[i]];
        }
        srsran_rm_turbo_rx_lut(e, &sb[(size_t)(n - j->first) * SB_STRIDE], L, (uint32_t)cb_idx, 0);
      }
    }
    free(T);
    free(e);
  } else {
    srsran_tdec_force_not_sb(&h);
  }
  pthread_barrier_wait(j->barrier);
  j->t_start = now_s();
  for (uint32_t n = j->first; n < j->last && !j->err; n++) {
    int16_t* in  = sb ? &sb[(size_t)(n - j->first) * SB_STRIDE] : &j->in[(size_t)n * L];
    uint8_t* out = &j->out[(size_t)n * (j->K / 8)];
    srsran_tdec_new_cb(&h, j->K);
    uint32_t noi = 0;
    int      ok  = 0;
    if (!j->early_stop) {
      srsran_tdec_run_all(&h, in, out, j->max_iter, j->K);
      noi = j->max_iter;
      ok  = srsran_crc_checksum_byte(&crc, out, j->K) == 0;
    } else {
      do {
        srsran_tdec_iteration(&h, in, out);
        noi++;
        if (!srsran_crc_checksum_byte(&crc, out, j->K) && noi >= 2) {
          ok = 1;
        }
      } while (noi < j->max_iter && !ok);
    }
    j->noi[n]    = (uint8_t)noi;
    j->crc_ok[n] = (uint8_t)ok;
  }
  j->t_end = now_s();
  srsran_tdec_free(&h);
  free(sb);
  return NULL;
}

/* returns wall seconds for decoding n code blocks with nthreads pthreads; pin: bit 0 = pin thread t to core t, bit 1 = feed the
 * decoder the production sub-block layout (see job_run) instead of natural-order input with srsran_tdec_force_not_sb */
double ref_tdec_batch(int       impl,
                      uint32_t  K,
                      int16_t*  in,
                      uint32_t  n,
                      uint32_t  max_iter,
                      int       early_stop,
                      int       nthreads,
                      int       pin,
                      uint8_t*  out,
                      uint8_t*  noi,
                      uint8_t*  crc_ok)
{
  ref_init();
  if (nthreads < 1) {
    nthreads = 1;
  }
  pthread_t*        th   = calloc(nthreads, sizeof(pthread_t));
  job_t*            jobs = calloc(nthreads, sizeof(job_t));
  pthread_barrier_t barrier;
  pthread_barrier_init(&barrier, NULL, nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (job_t){impl, K, in, out, noi, crc_ok, (uint32_t)((uint64_t)n * t / nthreads),
                      (uint32_t)((uint64_t)n * (t + 1) / nthreads), max_iter, early_stop, (pin & 1) ? t : -1, 0, 0, 0, &barrier, (pin & 2) ? 1 : 0};
    pthread_create(&th[t], NULL, job_run, &jobs[t]);
  }
  int    err = 0;
  double t0 = 1e300, t1 = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    err |= jobs[t].err;
    if (jobs[t].t_start < t0) t0 = jobs[t].t_start;
    if (jobs[t].t_end > t1) t1 = jobs[t].t_end;
  }
  pthread_barrier_destroy(&barrier);
  free(th);
  free(jobs);
  /* wall time from the first thread leaving the start barrier to the last thread finishing its share */
  return err ? -1.0 : t1 - t0;
}

/*
 * The reference's encode_tb_off loop (lib/src/phy/phch/sch.c:240-358; static there and sch.c drags in the whole PHY, so
 * the loop is restated here around the *reference's own* primitives): srsran_tcod_encode_lut (CRC attach + turbo
 * encode on packed bytes) and srsran_rm_turbo_tx_lut (interleave into the circular buffer at rv 0, bit selection).
 * Each CB's circular buffer is primed with an rv=0 call when rv != 0, as a HARQ process would have done.
 * e_bits must hold (nof_e_bits+7)/8 + 8 bytes; it is zeroed first.
 */
int ref_encode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const uint8_t* data, uint8_t* e_bits)
{
  ref_init();
  if (!data || !e_bits) {
    return -2;
  }
  srsran_cbsegm_t seg;
  if (srsran_cbsegm(&seg, tbs)) {
    return -1;
  }
  if (seg.F || Qm == 0 || rv >= 4) {
    return -1;
  }
  static srsran_tcod_t enc;
  static int           enc_ready = 0;
  pthread_mutex_lock(&g_lock);
  if (!enc_ready) {
    srsran_tcod_init(&enc, MAX_K);
    enc_ready = 1;
  }
  srsran_crc_t crc_tb, crc_cb;
  srsran_crc_init(&crc_tb, CRC24A, 24);
  srsran_crc_init(&crc_cb, CRC24B, 24);
  uint8_t* cb_in  = calloc(MAX_K / 8 + 16, 1);
  uint8_t* parity = calloc(3 * MAX_K / 8 + 16, 1);
  uint8_t* w_buff = calloc(SOFTBUFFER_SIZE, 1);
  uint8_t* scratch = calloc((nof_e_bits + 7) / 8 + 64, 1);
  int      ret = 0;
  memset(e_bits, 0, (nof_e_bits + 7) / 8);
  uint32_t Gp = nof_e_bits / Qm, gamma = seg.C > 0 ? Gp % seg.C : Gp;
  srsran_crc_set_init(&crc_tb, 0);
  uint32_t wp = 0, rp = 0;
  for (uint32_t i = 0; i < seg.C; i++) {
    uint32_t cb_len = i < seg.C2 ? seg.K2 : seg.K1, cblen_idx = i < seg.C2 ? seg.K2_idx : seg.K1_idx;
    uint32_t rlen   = seg.C > 1 ? cb_len - 24 : cb_len;
    uint32_t n_e    = (i <= seg.C - gamma - 1) ? Qm * (Gp / seg.C) : Qm * ((uint32_t)ceilf((float)Gp / seg.C));
    bool     last   = i == seg.C - 1;
    memcpy(cb_in, &data[rp / 8], (last ? rlen - 24 : rlen) / 8);
    srsran_tcod_encode_lut(&enc, &crc_tb, seg.C > 1 ? &crc_cb : NULL, cb_in, parity, cblen_idx, last);
    /* the circular buffer (softbuffer->buffer_b[i]) is filled only at rv 0 (rm_turbo.c:358-366): prime it first */
    if (rv != 0 && srsran_rm_turbo_tx_lut(w_buff, cb_in, parity, &scratch[wp / 8], cblen_idx, n_e, wp % 8, 0)) {
      ret = -1;
      break;
    }
    if (srsran_rm_turbo_tx_lut(w_buff, cb_in, parity, &e_bits[wp / 8], cblen_idx, n_e, wp % 8, rv)) {
      ret = -1;
      break;
    }
    rp += rlen;
    wp += n_e;
  }
  pthread_mutex_unlock(&g_lock);
  free(cb_in);
  free(parity);
  free(w_buff);
  free(scratch);
  return ret;
}

/* ------------------------------------------------------------------ the LITERAL sch.c entry points */
#include "srsran/phy/phch/pdsch_cfg.h"
#include "srsran/phy/phch/sch.h"

static srsran_sch_t g_sch;
static int          g_sch_ready = 0;
static int          sch_ready(void)
{
  if (!g_sch_ready) {
    if (srsran_sch_init(&g_sch)) {
      return -1;
    }
    g_sch_ready = 1;
  }
  return 0;
}
static srsran_mod_t mod_of(uint32_t Qm)
{
  switch (Qm) {
    case 1:
      return SRSRAN_MOD_BPSK;
    case 2:
      return SRSRAN_MOD_QPSK;
    case 4:
      return SRSRAN_MOD_16QAM;
    case 6:
      return SRSRAN_MOD_64QAM;
    default:
      return SRSRAN_MOD_256QAM;
  }
}

/*
 * srsran_dlsch_encode2 itself (sch.c:618-650 -> static encode_tb -> encode_tb_off): the reference's own transport-block
 * encoder, untouched. A fresh tx soft buffer per call; for rv != 0 the same buffer first sees the rv = 0 transmission,
 * as a HARQ process does. e_bits must hold (nof_e_bits+7)/8 + 8 bytes.
 */
int ref_dlsch_encode(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, uint8_t* data, uint8_t* e_bits)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_tx_t sb;
    if (srsran_softbuffer_tx_init(&sb, 110) == 0) {
      srsran_pdsch_cfg_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.grant.nof_tb         = 1;
      cfg.grant.tb[0].tbs      = (int)tbs;
      cfg.grant.tb[0].mod      = mod_of(Qm);
      cfg.grant.tb[0].nof_bits = nof_e_bits;
      cfg.grant.tb[0].enabled  = true;
      cfg.softbuffers.tx[0]    = &sb;
      ret                      = 0;
      if (rv != 0) {
        uint8_t* scratch    = calloc((nof_e_bits + 7) / 8 + 64, 1);
        cfg.grant.tb[0].rv  = 0;
        ret                 = srsran_dlsch_encode2(&g_sch, &cfg, data, scratch, 0, 1);
        free(scratch);
      }
      if (ret == 0) {
        cfg.grant.tb[0].rv = (int)rv;
        ret                = srsran_dlsch_encode2(&g_sch, &cfg, data, e_bits, 0, 1);
      }
      srsran_softbuffer_tx_free(&sb);
    }
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}

/*
 * srsran_dlsch_decode2 itself (sch.c:580-609 -> static decode_tb -> decode_tb_cb) with the PRODUCTION decoder
 * (srsran_tdec_init = AUTO = AVX2/SSE windowed, sub-block input layout). Its soft values and iteration counts are not
 * the parity target, but everything the loop decides from CRCs is: return code, data bytes and their
 * layout, cb_crc / tb_crc flags, cached code blocks across HARQ transmissions. Used to cross-check the restated loop.
 * The soft buffer lives across calls in *handle (create with ref_dlsch_rx_new, free with ref_dlsch_rx_free).
 */
void* ref_dlsch_rx_new(void)
{
  srsran_softbuffer_rx_t* sb = calloc(1, sizeof(srsran_softbuffer_rx_t));
  if (sb && srsran_softbuffer_rx_init(sb, 110)) {
    free(sb);
    sb = NULL;
  }
  return sb;
}
/* a soft buffer with room for max_cb code blocks (srsran_softbuffer_rx_init_guru, softbuffer.c:48): the 110-PRB one above holds 16,
 * a two-layer transport block of 100 PRB has 25 */
void* ref_dlsch_rx_new_guru(uint32_t max_cb)
{
  srsran_softbuffer_rx_t* sb = calloc(1, sizeof(srsran_softbuffer_rx_t));
  if (sb && srsran_softbuffer_rx_init_guru(sb, max_cb, SOFTBUFFER_SIZE)) {
    free(sb);
    sb = NULL;
  }
  return sb;
}
void ref_dlsch_rx_free(void* h)
{
  if (h) {
    srsran_softbuffer_rx_free((srsran_softbuffer_rx_t*)h);
    free(h);
  }
}
void ref_dlsch_rx_reset(void* h, uint32_t tbs)
{
  srsran_softbuffer_rx_reset_tbs((srsran_softbuffer_rx_t*)h, tbs);
}
int ref_dlsch_decode(void* h, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, int16_t* e_bits, uint32_t max_iterations, uint8_t* data,
                     uint8_t* cb_crc, uint8_t* tb_crc, float* avg_iterations)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_rx_t* sb = (srsran_softbuffer_rx_t*)h;
    srsran_pdsch_cfg_t      cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.grant.nof_tb         = 1;
    cfg.grant.tb[0].tbs      = (int)tbs;
    cfg.grant.tb[0].mod      = mod_of(Qm);
    cfg.grant.tb[0].rv       = (int)rv;
    cfg.grant.tb[0].nof_bits = nof_e_bits;
    cfg.grant.tb[0].enabled  = true;
    cfg.softbuffers.rx[0]    = sb;
    srsran_sch_set_max_noi(&g_sch, max_iterations);
    ret = srsran_dlsch_decode2(&g_sch, &cfg, e_bits, data, 0, 1);
    for (uint32_t i = 0; i < sb->max_cb; i++) {
      cb_crc[i] = sb->cb_crc[i] ? 1 : 0;
    }
    *tb_crc         = sb->tb_crc ? 1 : 0;
    *avg_iterations = srsran_sch_last_noi(&g_sch);
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}
/*
 * The multi-codeword / multi-layer forms of the two calls above (sch.c:580-609, 618-650): grant with nof_tb transport blocks,
 * the call addresses codeword tb_idx and passes nof_layers; Nl = 2 whenever nof_layers != nof_tb, and decode_tb / encode_tb
 * then see Qm * Nl (BASELINE config 3: 64QAM on 2 layers => 12). The other codeword of a 2-TB grant is filled with a
 * different TBS / modulation on purpose: the call must not look at it.
 */
static void cw_cfg_fill(srsran_pdsch_cfg_t* cfg, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, uint32_t tb_idx, uint32_t nof_tb)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->grant.nof_tb = nof_tb;
  for (uint32_t t = 0; t < nof_tb && t < SRSRAN_MAX_CODEWORDS; t++) {
    cfg->grant.tb[t].tbs      = (t == tb_idx) ? (int)tbs : 1000;
    cfg->grant.tb[t].mod      = (t == tb_idx) ? mod_of(Qm) : SRSRAN_MOD_QPSK;
    cfg->grant.tb[t].rv       = (t == tb_idx) ? (int)rv : 1;
    cfg->grant.tb[t].nof_bits = (t == tb_idx) ? nof_e_bits : 2400;
    cfg->grant.tb[t].enabled  = true;
  }
}
int ref_dlsch_decode_cw(void* h, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, int16_t* e_bits, uint32_t max_iterations,
                        uint8_t* data, uint8_t* cb_crc, uint8_t* tb_crc, float* avg_iterations, uint32_t tb_idx, uint32_t nof_layers,
                        uint32_t nof_tb)
{
  ref_init();
  if (tb_idx >= SRSRAN_MAX_CODEWORDS || tb_idx >= nof_tb) {
    return -2;
  }
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_rx_t* sb = (srsran_softbuffer_rx_t*)h;
    srsran_pdsch_cfg_t      cfg;
    cw_cfg_fill(&cfg, tbs, Qm, rv, nof_e_bits, tb_idx, nof_tb);
    cfg.softbuffers.rx[tb_idx] = sb;
    srsran_sch_set_max_noi(&g_sch, max_iterations);
    ret = srsran_dlsch_decode2(&g_sch, &cfg, e_bits, data, (int)tb_idx, nof_layers);
    for (uint32_t i = 0; i < sb->max_cb; i++) {
      cb_crc[i] = sb->cb_crc[i] ? 1 : 0;
    }
    *tb_crc         = sb->tb_crc ? 1 : 0;
    *avg_iterations = srsran_sch_last_noi(&g_sch);
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}
int ref_dlsch_encode_cw(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, uint8_t* data, uint8_t* e_bits, uint32_t tb_idx,
                        uint32_t nof_layers, uint32_t nof_tb)
{
  ref_init();
  if (tb_idx >= SRSRAN_MAX_CODEWORDS || tb_idx >= nof_tb) {
    return -2;
  }
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_tx_t sb;
    if (srsran_softbuffer_tx_init_guru(&sb, 32, SOFTBUFFER_SIZE) == 0) {  /* room for the 25 code blocks of a two-layer TB */
      srsran_pdsch_cfg_t cfg;
      cw_cfg_fill(&cfg, tbs, Qm, 0, nof_e_bits, tb_idx, nof_tb);
      cfg.softbuffers.tx[tb_idx] = &sb;
      ret                        = 0;
      if (rv != 0) {
        uint8_t* scratch = calloc((nof_e_bits + 7) / 8 + 64, 1);
        ret              = srsran_dlsch_encode2(&g_sch, &cfg, data, scratch, (int)tb_idx, nof_layers);
        free(scratch);
      }
      if (ret == 0) {
        cfg.grant.tb[tb_idx].rv = (int)rv;
        ret                     = srsran_dlsch_encode2(&g_sch, &cfg, data, e_bits, (int)tb_idx, nof_layers);
      }
      srsran_softbuffer_tx_free(&sb);
    }
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}

/* HARQ retransmission WITHOUT payload (sch.c:305 `if (data)`): rv 0 with data on a fresh soft buffer, then redundancy version rv
 * with data == NULL on the same soft buffer; e_bits receives the second transmission */
int ref_dlsch_encode_retx_null(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, uint8_t* data, uint8_t* e_bits)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_tx_t sb;
    if (srsran_softbuffer_tx_init(&sb, 110) == 0) {
      srsran_pdsch_cfg_t cfg;
      cw_cfg_fill(&cfg, tbs, Qm, 0, nof_e_bits, 0, 1);
      cfg.softbuffers.tx[0] = &sb;
      uint8_t* scratch      = calloc((nof_e_bits + 7) / 8 + 64, 1);
      ret                   = srsran_dlsch_encode2(&g_sch, &cfg, data, scratch, 0, 1);
      free(scratch);
      if (ret == 0) {
        cfg.grant.tb[0].rv = (int)rv;
        ret                = srsran_dlsch_encode2(&g_sch, &cfg, NULL, e_bits, 0, 1);
      }
      srsran_softbuffer_tx_free(&sb);
    }
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}

uint32_t ref_dlsch_rx_max_cb(void* h)
{
  return ((srsran_softbuffer_rx_t*)h)->max_cb;
}

/* ulsch_deinterleave itself (sch.c:994-1021; non-static, no header) */
void ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits, srsran_uci_bit_t* ri_bits,
                        uint32_t nof_ri_bits, uint8_t* ri_present, uint32_t* inteleaver_lut);
int ref_ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits, const uint32_t* ri_positions,
                           uint32_t nof_ri_bits)
{
  uint32_t          n   = H_prime_total * Qm;
  srsran_uci_bit_t* ri  = calloc(nof_ri_bits + 1, sizeof(srsran_uci_bit_t));
  uint8_t*          prs = calloc(n + 64, 1);
  uint32_t*         lut = calloc(n + 64, sizeof(uint32_t));
  for (uint32_t i = 0; i < nof_ri_bits; i++) {
    ri[i].position = ri_positions[i];
  }
  ulsch_deinterleave(q_bits, Qm, H_prime_total, N_pusch_symbs, g_bits, ri, nof_ri_bits, prs, lut);
  free(ri);
  free(prs);
  free(lut);
  return 0;
}

/* ------------------------------------------------------------------ srsran_ulsch_encode / srsran_ulsch_decode themselves */
#include "srsran/phy/phch/pusch_cfg.h"
static void pusch_cfg_fill(srsran_pusch_cfg_t* cfg, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_symb, uint32_t L_prb, uint32_t ri_len)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->grant.L_prb         = L_prb;
  cfg->grant.nof_symb      = nof_symb;
  cfg->grant.nof_re        = L_prb * 12 * nof_symb;
  cfg->grant.tb.tbs        = (int)tbs;
  cfg->grant.tb.mod        = mod_of(Qm);
  cfg->grant.tb.rv         = (int)rv;
  cfg->grant.tb.nof_bits   = cfg->grant.nof_re * Qm;
  cfg->grant.tb.enabled    = true;
  cfg->grant.last_tb       = cfg->grant.tb;
  cfg->uci_cfg.cqi.ri_len  = ri_len; /* 1-bit RI multiplexed on the PUSCH (no CQI, no ACK) */
  cfg->uci_offset.I_offset_ri  = 5;
  cfg->uci_offset.I_offset_cqi = 6;
  cfg->uci_offset.I_offset_ack = 5;
}

/* q_bits: packed interleaved bits, L_prb*12*nof_symb*Qm/8 bytes (+8 spare) */
int ref_ulsch_encode(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_symb, uint32_t L_prb, uint32_t ri_len, uint32_t ri_value, uint8_t* data,
                     uint8_t* q_bits)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_tx_t sb;
    if (srsran_softbuffer_tx_init(&sb, 110) == 0) {
      srsran_pusch_cfg_t cfg;
      pusch_cfg_fill(&cfg, tbs, Qm, rv, nof_symb, L_prb, ri_len);
      cfg.softbuffers.tx = &sb;
      srsran_uci_value_t uci;
      memset(&uci, 0, sizeof(uci));
      uci.ri          = (uint8_t)ri_value;
      uint8_t* g_bits = calloc(cfg.grant.tb.nof_bits / 8 + 64, 1);
      ret             = srsran_ulsch_encode(&g_sch, &cfg, data, &uci, g_bits, q_bits);
      free(g_bits);
      srsran_softbuffer_tx_free(&sb);
    }
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}

/* q_llr: L_prb*12*nof_symb*Qm interleaved LLRs (unscrambled); soft buffer handle as for ref_dlsch_decode */
int ref_ulsch_decode(void* h, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_symb, uint32_t L_prb, uint32_t ri_len, int16_t* q_llr,
                     uint32_t max_iterations, uint8_t* data, uint8_t* cb_crc, uint8_t* tb_crc, float* avg_iterations, uint8_t* ri_out)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_rx_t* sb = (srsran_softbuffer_rx_t*)h;
    srsran_pusch_cfg_t      cfg;
    pusch_cfg_fill(&cfg, tbs, Qm, rv, nof_symb, L_prb, ri_len);
    cfg.softbuffers.rx = sb;
    srsran_uci_value_t uci;
    memset(&uci, 0, sizeof(uci));
    uint32_t nb    = cfg.grant.tb.nof_bits;
    int16_t* g     = calloc(nb + 64, sizeof(int16_t));
    uint8_t* c_seq = calloc(nb + 64, 1);
    srsran_sch_set_max_noi(&g_sch, max_iterations);
    ret = srsran_ulsch_decode(&g_sch, &cfg, q_llr, g, c_seq, data, &uci);
    for (uint32_t i = 0; i < sb->max_cb; i++) {
      cb_crc[i] = sb->cb_crc[i] ? 1 : 0;
    }
    *tb_crc         = sb->tb_crc ? 1 : 0;
    *avg_iterations = srsran_sch_last_noi(&g_sch);
    *ri_out         = uci.ri;
    free(g);
    free(c_seq);
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}

/* srsran_dlsch_decode2 in 8-bit LLR mode (q->llr_is_8bit: rate de-matching and turbo decoding on int8 values) */
static float g_last_avg_noi = 0;
int ref_dlsch_decode8(void* h, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, int8_t* e_bits, uint32_t max_iterations, uint8_t* data,
                      uint8_t* cb_crc, uint8_t* tb_crc)
{
  ref_init();
  pthread_mutex_lock(&g_lock);
  int ret = -1;
  if (sch_ready() == 0) {
    srsran_softbuffer_rx_t* sb = (srsran_softbuffer_rx_t*)h;
    srsran_pdsch_cfg_t      cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.grant.nof_tb         = 1;
    cfg.grant.tb[0].tbs      = (int)tbs;
    cfg.grant.tb[0].mod      = mod_of(Qm);
    cfg.grant.tb[0].rv       = (int)rv;
    cfg.grant.tb[0].nof_bits = nof_e_bits;
    cfg.grant.tb[0].enabled  = true;
    cfg.softbuffers.rx[0]    = sb;
    srsran_sch_set_max_noi(&g_sch, max_iterations);
    g_sch.llr_is_8bit = true;
    ret               = srsran_dlsch_decode2(&g_sch, &cfg, (int16_t*)e_bits, data, 0, 1);
    g_sch.llr_is_8bit = false;
    for (uint32_t i = 0; i < sb->max_cb; i++) {
      cb_crc[i] = sb->cb_crc[i] ? 1 : 0;
    }
    *tb_crc = sb->tb_crc ? 1 : 0;
    g_last_avg_noi = srsran_sch_last_noi(&g_sch);
  }
  pthread_mutex_unlock(&g_lock);
  return ret;
}
/* srsran_sch_last_noi after the latest ref_dlsch_decode8 (kept out of its signature, which the drop-in tests bind) */
float ref_last_avg_iterations(void)
{
  return g_last_avg_noi;
}

/* srsran_sequence_apply_s itself (lib/src/phy/common/sequence.c:507-561) */
#include "srsran/phy/common/sequence.h"
void ref_sequence_apply_s(const int16_t* in, int16_t* out, uint32_t len, uint32_t c_init)
{
  srsran_sequence_apply_s(in, out, len, c_init);
}

/* srsran_demod_soft_demodulate_s itself (lib/src/phy/modem/demod_soft.c:871-894); symbols must be 32-byte aligned (_mm_load_ps) */
#include "srsran/phy/modem/demod_soft.h"
int ref_demod_soft_demodulate_s(int mod, const float* symbols, int16_t* llr, int nsymbols)
{
  cf_t*    s = srsran_vec_cf_malloc(nsymbols + 8);
  int16_t* l = srsran_vec_i16_malloc(8 * nsymbols + 64);
  memcpy(s, symbols, sizeof(cf_t) * nsymbols);
  int r = srsran_demod_soft_demodulate_s((srsran_mod_t)mod, s, l, nsymbols);
  static const int bps[5] = {1, 2, 4, 6, 8};
  if (r == 0) {
    memcpy(llr, l, sizeof(int16_t) * (size_t)bps[mod] * nsymbols);
  }
  free(s);
  free(l);
  return r;
}
