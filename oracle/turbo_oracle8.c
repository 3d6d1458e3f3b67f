/*
 * turbo_oracle8.c - TEST INFRASTRUCTURE ONLY (part of oracle/libturbo_oracle.so; never linked into the product).
 *
 * Clean-room CPU restatement of the reference's 8-BIT LLR mode: what srsran_tdec_iteration_8bit /
 * srsran_tdec_run_all_8bit (lib/src/phy/fec/turbo/turbodecoder.c:458-484,551-577) compute in AUTO mode for the block sizes
 * that really run in 8-bit arithmetic, and srsran_rm_turbo_rx_lut_8bit (lib/src/phy/fec/turbo/rm_turbo.c:447-483).
 *
 * The 8-bit decoders are NOT the generic algorithm in fewer bits: they are the windowed SIMD decoders of
 * lib/include/srsran/phy/fec/turbo/turbodecoder_win.h instantiated with llr_t = int8_t (:180-186, :217-283):
 *   - the code block is cut into NW windows of S = K / NW steps (NW = 32 for K > 2048 and K % 32 == 0, else 16 for K > 800 and
 *     K % 16 == 0: srsran_tdec_autoimp_get_subblocks_8bit, turbodecoder.c:410-424); every window runs its own recursions;
 *   - a window's backward recursion starts from the state its RIGHT neighbour reaches after a warm-up over that neighbour's
 *     first 40 steps (win_overlap_len) from the all-"unknown" state, the last window from the three termination steps;
 *     the forward recursion likewise from the LEFT neighbour's last 40 steps, the first window from the known state (:551-651,
 *     :654-800). "Unknown" and the known start are both the all-zero vector, because INF is 0 in the 8-bit instantiation;
 *   - every add / subtract saturates to [-128, 127] (_mm256_adds_epi8 / _mm256_subs_epi8); the termination steps use the scalar
 *     helper that saturates only upwards and wraps downwards (:469-477);
 *   - the state vector is normalised by its MAXIMUM after every step except the one with loop index 0 (:479-497);
 *   - the extrinsic output is (max1 - max0) >> 1, arithmetic shift per 8-bit element (:761-766).
 * Around it the turbo schedule of turbodecoder_iter.h:72-144 in 8 bits: srsran_vec_sub_bbb saturates - except, in the AVX2
 * build the oracle is pinned to, on the last K % 32 elements OF THE WINDOW-INTERLEAVED ARRAY, which its scalar tail subtracts
 * with wrap-around (vector_simd.c:162-190); only K = 16 mod 32 has such a tail and it holds the last step of every window.
 *
 * Everything here is written in NATURAL trellis order (element n = window * S + step): the reference's window-interleaved
 * storage is a layout, not arithmetic. Sizes outside the two window decoders (K <= 800, or K not a multiple of 16) are not
 * 8-bit arithmetic in the reference either (it widens to its SSE int16 decoders, turbodecoder.c:443-476): orc_tdec8_windows
 * returns 0 for them and the functions below refuse them.
 *
 * Parity status: PINNED against the compiled reference (oracle/_ref, AVX2 build): tests/test_oracle8_vs_ref.py.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "turbo_oracle.h"

#define OVERLAP 40

static inline int8_t sat8(int v)
{
  return (int8_t)(v > 127 ? 127 : (v < -128 ? -128 : v));
}
static inline int8_t adds8(int8_t a, int8_t b)
{
  return sat8((int)a + (int)b);
}
static inline int8_t subs8(int8_t a, int8_t b)
{
  return sat8((int)a - (int)b);
}
/* the scalar helper of the termination steps: saturates at +127 only, wraps below -128 (turbodecoder_win.h:469-477) */
static inline int8_t sadd_tail(int8_t a, int8_t b)
{
  int z = (int)a + (int)b;
  return z > 127 ? 127 : (int8_t)z;
}
static inline int8_t max8(int8_t a, int8_t b)
{
  return a > b ? a : b;
}

uint32_t orc_tdec8_windows(uint32_t K)
{
  if (orc_cbindex_exact(K) < 0) {
    return 0;
  }
  if (K % 32 == 0 && K > 2048) {
    return 32;
  }
  if (K % 16 == 0 && K > 800) {
    return 16;
  }
  return 0;
}

/* one backward step: o <- beta at the previous trellis position (saturating), branch metrics from x (systematic + a-priori), y */
static void bstep(int8_t o[8], int8_t x, int8_t y)
{
  int8_t xy = adds8(x, y);
  int8_t a[8] = {adds8(o[4], xy), o[4], adds8(o[5], y), adds8(o[5], x), adds8(o[6], x), adds8(o[6], y), o[7], adds8(o[7], xy)};
  int8_t b[8] = {o[0], adds8(o[0], xy), adds8(o[1], x), adds8(o[1], y), adds8(o[2], y), adds8(o[2], x), adds8(o[3], xy), o[3]};
  for (int i = 0; i < 8; i++) {
    o[i] = max8(a[i], b[i]);
  }
}
static void norm_max(int8_t o[8])
{
  int8_t m = o[0];
  for (int i = 1; i < 8; i++) {
    m = max8(m, o[i]);
  }
  for (int i = 0; i < 8; i++) {
    o[i] = subs8(o[i], m);
  }
}
/* branch sums of one forward step: z[i] = information bit 0 into state i, w[i] = information bit 1 into state i */
static void abranches(const int8_t o[8], int8_t x, int8_t y, int8_t z[8], int8_t w[8])
{
  int8_t xy = adds8(x, y);
  z[0] = o[0]; z[1] = adds8(o[3], y); z[2] = adds8(o[4], y); z[3] = o[7];
  z[4] = o[1]; z[5] = adds8(o[2], y); z[6] = adds8(o[5], y); z[7] = o[6];
  w[0] = adds8(o[1], xy); w[1] = adds8(o[2], x); w[2] = adds8(o[5], x); w[3] = adds8(o[6], xy);
  w[4] = adds8(o[0], xy); w[5] = adds8(o[3], x); w[6] = adds8(o[4], x); w[7] = adds8(o[7], xy);
}

typedef struct {
  uint32_t  K, NW, S, n_iter;
  int8_t *  syst, *par0, *par1, *app1, *app2, *ext1, *ext2, *x, *beta, *mem;
  uint16_t *fwd, *rev;
} tdec8_t;

static int tdec8_open(tdec8_t* d, uint32_t K)
{
  memset(d, 0, sizeof(*d));
  d->NW = orc_tdec8_windows(K);
  if (!d->NW) {
    return -1;
  }
  d->K     = K;
  d->S     = K / d->NW;
  size_t n = K + 16;
  d->mem   = calloc(8 * n + 8 * (size_t)(K + d->NW), 1);
  d->syst  = d->mem;
  d->par0  = d->syst + n;
  d->par1  = d->par0 + n;
  d->app1  = d->par1 + n;
  d->app2  = d->app1 + n;
  d->ext1  = d->app2 + n;
  d->ext2  = d->ext1 + n;
  d->x     = d->ext2 + n;
  d->beta  = d->x + n; /* [window][step 0..S][8] */
  d->fwd   = malloc(2 * K * sizeof(uint16_t));
  d->rev   = d->fwd + K;
  orc_qpp(K, d->fwd, d->rev);
  return 0;
}
static void tdec8_close(tdec8_t* d)
{
  free(d->mem);
  free(d->fwd);
}

/* one constituent decode: input[K+3] (systematic, or the a-priori stream for the second decoder), app[K] or NULL, parity[K+3] */
static void win_decode(tdec8_t* d, const int8_t* input, const int8_t* app, const int8_t* parity, int8_t* out)
{
  const uint32_t K = d->K, NW = d->NW, S = d->S;
  int8_t*        x = d->x;
  for (uint32_t n = 0; n < K; n++) {
    x[n] = app ? adds8(app[n], input[n]) : input[n]; /* simd_add(ap, x), turbodecoder_win.h:608-611 */
  }
  /* ---- backward: warm-up of every window over ITS first 40 steps, handed to the window on its left */
  int8_t init[32][8];
  for (uint32_t w = 0; w + 1 < NW; w++) {
    int8_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* simd_set1(-INF), INF = 0 */
    for (int k = OVERLAP - 1; k >= 0; k--) {
      bstep(o, x[(w + 1) * S + (uint32_t)k], parity[(w + 1) * S + (uint32_t)k]);
      if (k != 0) {
        norm_max(o);
      }
    }
    memcpy(init[w], o, 8);
  }
  {
    /* last window: the three termination steps from {0, -INF x 7} = all zero, scalar helper, no normalisation (:499-549) */
    int8_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = (int)K + 2; k >= (int)K; k--) {
      int8_t xv = input[k], yv = parity[k]; /* (no a-priori on the termination steps) */
      int8_t xy = sadd_tail(xv, yv);
      int8_t a[8] = {sadd_tail(o[4], xy), o[4], sadd_tail(o[5], yv), sadd_tail(o[5], xv), sadd_tail(o[6], xv), sadd_tail(o[6], yv), o[7], sadd_tail(o[7], xy)};
      int8_t b[8] = {o[0], sadd_tail(o[0], xy), sadd_tail(o[1], xv), sadd_tail(o[1], yv), sadd_tail(o[2], yv), sadd_tail(o[2], xv), sadd_tail(o[3], xy), o[3]};
      for (int i = 0; i < 8; i++) {
        o[i] = max8(a[i], b[i]);
      }
    }
    memcpy(init[NW - 1], o, 8);
  }
  /* ---- backward, main pass: beta[w][k] stored BEFORE the normalisation, beta[w][S] = the start state */
  for (uint32_t w = 0; w < NW; w++) {
    int8_t* B = &d->beta[(size_t)w * (S + 1) * 8];
    int8_t  o[8];
    memcpy(o, init[w], 8);
    memcpy(&B[S * 8], o, 8);
    for (int k = (int)S - 1; k >= 0; k--) {
      bstep(o, x[w * S + (uint32_t)k], parity[w * S + (uint32_t)k]);
      memcpy(&B[(size_t)k * 8], o, 8);
      if (k != 0) {
        norm_max(o);
      }
    }
  }
  /* ---- forward: warm-up over the LAST 40 steps of every window, handed to the window on its right */
  int8_t z[8], wv[8];
  for (uint32_t w = 1; w < NW; w++) {
    int8_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t k = 0; k < OVERLAP; k++) {
      uint32_t n = (w - 1) * S + (S - OVERLAP) + k;
      abranches(o, x[n], parity[n], z, wv);
      for (int i = 0; i < 8; i++) {
        o[i] = max8(z[i], wv[i]);
      }
      if (k != 0) {
        norm_max(o);
      }
    }
    memcpy(init[w], o, 8);
  }
  memset(init[0], 0, 8); /* known start state {0, -INF x 7}, INF = 0 */
  /* ---- forward, main pass with the extrinsic output */
  for (uint32_t w = 0; w < NW; w++) {
    const int8_t* B = &d->beta[(size_t)w * (S + 1) * 8];
    int8_t        o[8];
    memcpy(o, init[w], 8);
    for (uint32_t k = 0; k < S; k++) {
      uint32_t n = w * S + k;
      abranches(o, x[n], parity[n], z, wv);
      const int8_t* b  = &B[(size_t)(k + 1) * 8];
      int8_t        m0 = adds8(b[0], z[0]), m1 = adds8(b[0], wv[0]);
      for (int i = 1; i < 8; i++) {
        m0 = max8(m0, adds8(b[i], z[i]));
        m1 = max8(m1, adds8(b[i], wv[i]));
      }
      int8_t l = subs8(m1, m0);
      out[n]   = (int8_t)(l >> 1); /* simd_rb_shift(out, 1): arithmetic shift of the 8-bit element */
      for (int i = 0; i < 8; i++) {
        o[i] = max8(z[i], wv[i]);
      }
      if (k != 0) {
        norm_max(o);
      }
    }
  }
}

/* srsran_vec_sub_bbb on the window-interleaved arrays (AVX2 build): saturating, except the last K % 32 interleaved elements */
static void sub_glue(const tdec8_t* d, const int8_t* a, const int8_t* b, int8_t* r)
{
  const uint32_t K = d->K, NW = d->NW, S = d->S, tail0 = K - K % 32;
  for (uint32_t n = 0; n < K; n++) {
    uint32_t idx = (n % S) * NW + n / S; /* position in the interleaved array */
    r[n]         = idx >= tail0 ? (int8_t)((int)a[n] - (int)b[n]) : subs8(a[n], b[n]);
  }
}

static void tdec8_half_iteration(tdec8_t* d, const int8_t* in)
{
  uint32_t K = d->K, n = d->n_iter;
  if (n == 0) {
    for (uint32_t i = 0; i < K; i++) {
      d->syst[i] = in[3 * i];
      d->par0[i] = in[3 * i + 1];
      d->par1[i] = in[3 * i + 2];
    }
    for (uint32_t j = 0; j < 3; j++) {
      d->syst[K + j] = in[3 * K + 2 * j];
      d->par0[K + j] = in[3 * K + 2 * j + 1];
      d->app2[K + j] = in[3 * K + 6 + 2 * j];
      d->par1[K + j] = in[3 * K + 6 + 2 * j + 1];
    }
  }
  if ((n & 1) == 0) {
    if (n) {
      sub_glue(d, d->app1, d->ext1, d->app1);
    }
    win_decode(d, d->syst, n ? d->app1 : NULL, d->par0, d->ext1);
  } else {
    if (n > 1) {
      sub_glue(d, d->ext1, d->app1, d->ext1);
    }
    for (uint32_t i = 0; i < K; i++) {
      d->app2[d->rev[i]] = d->ext1[i];
    }
    win_decode(d, d->app2, NULL, d->par1, d->ext2);
    for (uint32_t i = 0; i < K; i++) {
      d->app1[d->fwd[i]] = d->ext2[i];
    }
  }
  d->n_iter++;
}

static void tdec8_decide(const tdec8_t* d, uint8_t* out)
{
  const int8_t* src = (d->n_iter % 2) ? d->ext1 : d->app1;
  for (uint32_t i = 0; i < d->K / 8; i++) {
    uint8_t b = 0;
    for (int j = 0; j < 8; j++) {
      b = (uint8_t)((b << 1) | (src[8 * i + j] > 0 ? 1 : 0));
    }
    out[i] = b;
  }
}

/* hard decisions after every half-iteration; dump (optional) [it][3][K]: ext1, ext2, app1 */
int orc_tdec8_trace(uint32_t K, const int8_t* in, uint32_t nof_iter, uint8_t* out_bytes, int8_t* dump)
{
  tdec8_t d;
  if (tdec8_open(&d, K)) {
    return -1;
  }
  for (uint32_t it = 0; it < nof_iter; it++) {
    tdec8_half_iteration(&d, in);
    tdec8_decide(&d, &out_bytes[it * (K / 8)]);
    if (dump) {
      memcpy(&dump[(it * 3 + 0) * K], d.ext1, K);
      memcpy(&dump[(it * 3 + 1) * K], d.ext2, K);
      memcpy(&dump[(it * 3 + 2) * K], d.app1, K);
    }
  }
  tdec8_close(&d);
  return 0;
}

/* ------------------------------------------------------------------ batch with CRC24B early stop (sch.c:426-456 semantics) */
typedef struct {
  uint32_t      K, first, last, max_iter;
  int           early_stop;
  const int8_t* in;
  uint8_t *     out, *noi, *crc_ok;
} b8job_t;

static void* b8job_run(void* arg)
{
  b8job_t* j = (b8job_t*)arg;
  tdec8_t  d;
  if (tdec8_open(&d, j->K)) {
    return NULL;
  }
  uint32_t L = 3 * j->K + 12;
  for (uint32_t n = j->first; n < j->last; n++) {
    const int8_t* in  = &j->in[(size_t)n * L];
    uint8_t*      out = &j->out[(size_t)n * (j->K / 8)];
    d.n_iter          = 0;
    int      ok       = 0;
    uint32_t noi      = 0;
    do {
      tdec8_half_iteration(&d, in);
      noi++;
      if (j->early_stop) {
        tdec8_decide(&d, out);
        if (noi >= 2 && orc_crc_bytes(ORC_CRC24B, 24, out, (int)j->K) == 0) {
          ok = 1;
        }
      }
    } while (noi < j->max_iter && !ok);
    if (!j->early_stop) {
      tdec8_decide(&d, out);
      ok = orc_crc_bytes(ORC_CRC24B, 24, out, (int)j->K) == 0;
    }
    j->noi[n]    = (uint8_t)noi;
    j->crc_ok[n] = (uint8_t)ok;
  }
  tdec8_close(&d);
  return NULL;
}

double orc_tdec8_batch(uint32_t K, const int8_t* in, uint32_t n, uint32_t max_iter, int early_stop, int nthreads, uint8_t* out, uint8_t* noi,
                       uint8_t* crc_ok)
{
  if (!orc_tdec8_windows(K)) {
    return -1.0;
  }
  if (nthreads < 1) {
    nthreads = 1;
  }
  if (max_iter < 1) {
    max_iter = 1;
  }
  pthread_t*      th   = calloc((size_t)nthreads, sizeof(pthread_t));
  b8job_t*        jobs = calloc((size_t)nthreads, sizeof(b8job_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (b8job_t){K, (uint32_t)((uint64_t)n * t / nthreads), (uint32_t)((uint64_t)n * (t + 1) / nthreads), max_iter, early_stop, in, out, noi, crc_ok};
    pthread_create(&th[t], NULL, b8job_run, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ rate de-matching, 8 bit (rm_turbo.c:447-483, :586-667) */
/* output[T[i mod L]] += input[i], int8 WRAP-around (plain C `+=` on int8_t), natural layout */
int orc_rm_rx8(const int8_t* e, int8_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  if (cb_idx >= ORC_NOF_CB_SIZES || rv >= 4) {
    return -2;
  }
  uint32_t  L = 3 * (uint32_t)orc_cbsize(cb_idx) + 12;
  uint16_t* T = malloc(L * sizeof(uint16_t));
  orc_rm_table(cb_idx, rv, T);
  for (uint32_t i = 0; i < E; i++) {
    buf[T[i % L]] = (int8_t)((int)buf[T[i % L]] + (int)e[i]);
  }
  free(T);
  return 0;
}

/* ------------------------------------------------------------------ transport-block loop with q->llr_is_8bit (sch.c:371-494) */
/* Same loop as orc_decode_tb on int8 e-bits and int8 soft buffers (buffer_b[r][ORC_SOFTBUFFER_SIZE] bytes). Only transport
 * blocks whose code-block sizes run in the 8-bit window decoders (orc_tdec8_windows != 0) are accepted: -3 otherwise. */
int orc_decode_tb8(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const int8_t* e_bits, uint32_t max_iterations, int8_t* buffer_b,
                   uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc, uint8_t* data, uint32_t* cb_noi, float* avg_iterations)
{
  uint32_t sg[12];
  if (orc_cbsegm(tbs, sg)) {
    return -1;
  }
  if (Qm == 0 || !e_bits || !data) {
    return -2;
  }
  uint32_t F = sg[0], C = sg[1], K1 = sg[2], K2 = sg[3], K1i = sg[4], K2i = sg[5], C1 = sg[6];
  if (tbs == 0 || C == 0) {
    return 0;
  }
  if (F) {
    return -2;
  }
  if (!orc_tdec8_windows(K1) || (C1 < C && !orc_tdec8_windows(K2))) {
    return -3;
  }
  if (max_iterations == 0) {
    max_iterations = 10;
  }
  float avg = 0;
  for (uint32_t r = 0; r < C; r++) {
    uint32_t K    = r < C1 ? K1 : K2;
    uint32_t rlen = C == 1 ? K : K - 24;
    uint8_t* dst  = &data[r * rlen / 8];
    cb_noi[r]     = 0;
    if (cb_crc[r]) {
      memcpy(dst, &sb_data[(size_t)r * (ORC_SOFTBUFFER_SIZE / 8)], rlen / 8);
      continue;
    }
    uint32_t Ki = r < C1 ? K1i : K2i;
    uint32_t Gp = nof_e_bits / Qm, gamma = Gp % C, n_e = Qm * (Gp / C);
    uint32_t rp = r * n_e, E = n_e;
    if (r > C - gamma) {
      E  = n_e + Qm;
      rp = (C - gamma) * n_e + (r - (C - gamma)) * E;
    }
    int8_t* buf = &buffer_b[(size_t)r * ORC_SOFTBUFFER_SIZE];
    orc_rm_rx8(&e_bits[rp], buf, E, Ki, rv);
    tdec8_t d;
    tdec8_open(&d, K);
    uint32_t noi = 0;
    int      ok  = 0;
    do {
      tdec8_half_iteration(&d, buf);
      tdec8_decide(&d, dst);
      noi++;
      avg += 1.0f;
      uint32_t crc = C > 1 ? orc_crc_bytes(ORC_CRC24B, 24, dst, (int)K) : orc_crc_bytes(ORC_CRC24A, 24, dst, (int)(tbs + 24));
      if (crc == 0 && noi >= 2) {
        cb_crc[r] = 1;
        ok        = 1;
      }
    } while (noi < max_iterations && !ok);
    cb_noi[r] = noi;
    tdec8_close(&d);
  }
  int all_ok = 1;
  for (uint32_t r = 0; r < C && all_ok; r++) {
    all_ok = cb_crc[r] != 0;
  }
  *tb_crc = (uint8_t)all_ok;
  if (!all_ok) {
    for (uint32_t r = 0; r < C; r++) {
      if (cb_crc[r]) {
        uint32_t K    = r < C1 ? K1 : K2;
        uint32_t rlen = C == 1 ? K : K - 24;
        memcpy(&sb_data[(size_t)r * (ORC_SOFTBUFFER_SIZE / 8)], &data[r * rlen / 8], rlen / 8);
      }
    }
  }
  *avg_iterations = avg / (float)C;
  if (!all_ok) {
    return -1;
  }
  if (C == 1) {
    return 0;
  }
  if (orc_crc_bytes(ORC_CRC24A, 24, data, (int)(tbs + 24)) == 0) {
    return 0;
  }
  memset(cb_crc, 0, C);
  return -1;
}
