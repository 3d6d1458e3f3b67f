/*
 * turbo_oracle.c - TEST INFRASTRUCTURE ONLY. NOT part of the product; the product library never links or loads it.
 *
 * Clean-room CPU restatement of the srsRAN 4G LTE turbo-decode hot path, written from the algorithm description in
 * SURVEY.md Appendix A (own structure: table-driven trellis, circular-buffer walk for rate matching). It is the
 * portable bit-exact checker for the CUDA engine on the GPU box, where /root/reference does not exist.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py proves every function here bit-identical to the compiled
 * reference (oracle/_ref, built from /root/reference by oracle/Makefile) - all 188 block sizes x 4 rv for the rate
 * matching tables, MAP/turbo soft outputs per half-iteration incl. int16 overflow regimes, CRC known answers
 * (lib/src/phy/fec/test/crc_test.h:37-38), encoder known answer (lib/src/phy/fec/turbo/test/turbodecoder_test.h:70-125).
 * tests/golden/ holds fixtures produced by the compiled reference for the GPU box.
 *
 * Reference functions restated (file:line under /root/reference/lib):
 *   orc_map_gen        <- tdec_gen_dec / map_gen_beta / map_gen_alpha   src/phy/fec/turbo/turbodecoder_gen.c:58-236
 *   orc_tdec_*         <- run_tdec_iteration_16bit                      include/srsran/phy/fec/turbo/turbodecoder_iter.h:72-144
 *                         tdec_gen_extract_input / _decision_byte       src/phy/fec/turbo/turbodecoder_gen.c:238-277
 *                         srsran_tdec_iteration / _run_all              src/phy/fec/turbo/turbodecoder.c:527-549
 *   orc_qpp            <- srsran_tc_interl_LTE_gen_interl               src/phy/fec/turbo/tc_interl_lte.c:69-109
 *   orc_rm_table/_rx   <- srsran_rm_turbo_gentable_receive, rx_lut_     src/phy/fec/turbo/rm_turbo.c:175-248,390-445
 *   orc_crc_*          <- srsran_crc_checksum_byte / gen_crc_table      src/phy/fec/crc.c:30-48,147-161
 *   orc_cbsegm         <- srsran_cbsegm                                 src/phy/fec/cbsegm.c:62-117
 *   orc_decode_tb      <- decode_tb_cb + decode_tb                      src/phy/phch/sch.c:371-494,509-573
 *   orc_tcod_encode    <- srsran_tcod_encode                            src/phy/fec/turbo/turbocoder.c:77-185
 */
#define _GNU_SOURCE
#include "turbo_oracle.h"
#include "../include/lte_qpp_params.h"

#include <pthread.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NEG_INF (-10000)

static inline int16_t wrap16(int v)
{
  return (int16_t)(uint16_t)((unsigned)v & 0xffffu);
}
static inline int16_t add16(int16_t a, int16_t b)
{
  return wrap16((int)a + (int)b);
}
static inline int16_t sub16(int16_t a, int16_t b)
{
  return wrap16((int)a - (int)b);
}
static inline int16_t max16(int16_t a, int16_t b)
{
  return a > b ? a : b;
}

/* ------------------------------------------------------------------ sizes, QPP, segmentation */
int orc_cbsize(uint32_t idx)
{
  return idx < LTE_NOF_CB_SIZES ? (int)lte_qpp_params[idx].K : -1;
}

int orc_cbindex(uint32_t K)
{
  for (int i = 0; i < LTE_NOF_CB_SIZES; i++) {
    if (lte_qpp_params[i].K >= K) {
      return i;
    }
  }
  return -1;
}

int orc_cbindex_exact(uint32_t K)
{
  int i = orc_cbindex(K);
  return (i >= 0 && lte_qpp_params[i].K == K) ? i : -1;
}

int orc_qpp(uint32_t K, uint16_t* fwd, uint16_t* rev)
{
  int idx = orc_cbindex_exact(K);
  if (idx < 0) {
    return -1;
  }
  uint64_t f1 = lte_qpp_params[idx].f1, f2 = lte_qpp_params[idx].f2;
  for (uint64_t i = 0; i < K; i++) {
    uint32_t p = (uint32_t)((f1 * i + f2 * i * i) % K);
    fwd[i]     = (uint16_t)p;
    rev[p]     = (uint16_t)i;
  }
  return 0;
}

int orc_cbsegm(uint32_t tbs, uint32_t* o)
{
  memset(o, 0, 12 * sizeof(uint32_t));
  if (tbs == 0) {
    return 0;
  }
  uint32_t B = tbs + 24, C, Bp;
  if (B <= ORC_MAX_K) {
    C  = 1;
    Bp = B;
  } else {
    C  = (B + (ORC_MAX_K - 24) - 1) / (ORC_MAX_K - 24);
    Bp = B + 24 * C;
  }
  o[8]     = tbs;
  o[1]     = C;
  int idx1 = orc_cbindex((Bp - 1) / C + 1);
  if (idx1 < 0) {
    return -1;
  }
  uint32_t K1 = (uint32_t)orc_cbsize(idx1);
  o[2]        = K1;
  o[4]        = (uint32_t)idx1;
  if (C == 1) {
    o[6] = 1;
  } else {
    /* idx1 > 0 whenever C > 1 (K1 > 3000) */
    uint32_t K2 = (uint32_t)orc_cbsize(idx1 - 1);
    o[3]        = K2;
    o[5]        = (uint32_t)(idx1 - 1);
    o[7]        = (C * K1 - Bp) / (K1 - K2);
    o[6]        = C - o[7];
  }
  o[9]  = 24;
  o[10] = 24;
  o[0]  = o[6] * o[2] + o[7] * o[3] - Bp;
  return 0;
}

/* ------------------------------------------------------------------ CRC */
static void crc_table(uint32_t poly, int order, uint32_t* tab)
{
  /* orders used on this path are 24 (and 16/8 for the known-answer tests) */
  uint32_t top = 1u << (order - 1), mask = (order == 32) ? 0xffffffffu : ((1u << order) - 1u);
  for (uint32_t b = 0; b < 256; b++) {
    uint32_t r = (order >= 8) ? (b << (order - 8)) : (b >> (8 - order));
    if (order >= 8) {
      for (int j = 0; j < 8; j++) {
        r = (r & top) ? ((r << 1) ^ poly) : (r << 1);
      }
      tab[b] = r & mask;
    } else {
      tab[b] = 0; /* not needed */
    }
  }
}

uint32_t orc_crc_bytes(uint32_t poly, int order, const uint8_t* data, int nbits)
{
  uint32_t tab[256];
  crc_table(poly, order, tab);
  uint32_t mask = (1u << order) - 1u, crc = 0;
  for (int i = 0; i < nbits / 8; i++) {
    uint32_t idx = ((crc >> (order - 8)) & 0xffu) ^ data[i];
    crc          = ((crc << 8) ^ tab[idx]) & mask;
  }
  return crc;
}

uint32_t orc_crc_bits(uint32_t poly, int order, const uint8_t* bits, int nbits)
{
  uint32_t top = 1u << (order - 1), mask = (1u << order) - 1u, crc = 0;
  for (int i = 0; i < nbits; i++) {
    uint32_t fb = ((crc & top) ? 1u : 0u) ^ (bits[i] & 1u);
    crc         = (crc << 1) & mask;
    if (fb) {
      crc ^= poly & mask;
    }
  }
  return crc;
}

/* ------------------------------------------------------------------ encoder (test-vector generation) */
/* One constituent RSC step, g0 = 1+D^2+D^3 (feedback), g1 = 1+D+D^3; state = 3 delay cells */
static inline uint8_t rsc_step(uint8_t* d, uint8_t u, int terminate, uint8_t* sys_out)
{
  uint8_t fb = d[1] ^ d[2];
  uint8_t in = terminate ? fb : u; /* during termination the input equals the feedback so the cell input is 0 */
  if (sys_out) {
    *sys_out = in;
  }
  uint8_t a = in ^ fb;
  uint8_t z = a ^ d[0] ^ d[2];
  d[2]      = d[1];
  d[1]      = d[0];
  d[0]      = a;
  return z;
}

int orc_tcod_encode(const uint8_t* in, uint8_t* out, uint32_t K)
{
  int idx = orc_cbindex_exact(K);
  if (idx < 0) {
    return -1;
  }
  uint16_t* fwd = malloc(2 * K * sizeof(uint16_t));
  orc_qpp(K, fwd, fwd + K);
  uint8_t d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0};
  for (uint32_t i = 0; i < K; i++) {
    out[3 * i]     = in[i] & 1;
    out[3 * i + 1] = rsc_step(d1, in[i] & 1, 0, NULL);
    out[3 * i + 2] = rsc_step(d2, in[fwd[i]] & 1, 0, NULL);
  }
  uint8_t* t = &out[3 * K];
  for (int j = 0; j < 3; j++) {
    uint8_t x;
    uint8_t z    = rsc_step(d1, 0, 1, &x);
    t[2 * j]     = x;
    t[2 * j + 1] = z;
  }
  for (int j = 0; j < 3; j++) {
    uint8_t x;
    uint8_t z        = rsc_step(d2, 0, 1, &x);
    t[6 + 2 * j]     = x;
    t[6 + 2 * j + 1] = z;
  }
  free(fwd);
  return 0;
}

/* ------------------------------------------------------------------ rate matching (36.212 5.1.4.1) */
static const uint8_t COLPERM[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                    1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/*
 * T[n] = index into the natural coded stream (3i+s, tails flattened as 3K..3K+11) of the n-th transmitted bit of
 * redundancy version rv, n < 3K+12. Built by placing every coded bit at its circular-buffer position and walking the
 * buffer from k0 skipping the <NULL> positions.
 */
int orc_rm_table(uint32_t cb_idx, uint32_t rv, uint16_t* T)
{
  int K = orc_cbsize(cb_idx);
  if (K < 0 || rv >= 4) {
    return -2;
  }
  uint32_t D = (uint32_t)K + 4, R = (D - 1) / 32 + 1, Kp = 32 * R, Nd = Kp - D, Ncb = 3 * Kp, L = 3 * D;
  int32_t* w = malloc(Ncb * sizeof(int32_t));
  for (uint32_t i = 0; i < Ncb; i++) {
    w[i] = -1;
  }
  for (uint32_t i = 0; i < D; i++) {
    uint32_t y = Nd + i;
    uint32_t v = COLPERM[y % 32] * R + y / 32; /* streams 0,1: column-permuted read-out position */
    w[v]           = (int32_t)(3 * i);
    w[Kp + 2 * v]  = (int32_t)(3 * i + 1);
    uint32_t t     = (y + Kp - 1) % Kp; /* stream 2: one-position cyclic shift before the same permutation */
    uint32_t v2    = COLPERM[t % 32] * R + t / 32;
    w[Kp + 2 * v2 + 1] = (int32_t)(3 * i + 2);
  }
  uint32_t k0 = R * (24 * rv + 2);
  uint32_t n = 0, j = 0;
  while (n < L) {
    int32_t src = w[(k0 + j) % Ncb];
    if (src >= 0) {
      T[n++] = (uint16_t)src;
    }
    j++;
  }
  free(w);
  return 0;
}

int orc_rm_rx(const int16_t* e, int16_t* buf, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  int K = orc_cbsize(cb_idx);
  if (K < 0 || rv >= 4) {
    return -2;
  }
  uint32_t  L = 3 * (uint32_t)K + 12;
  uint16_t* T = malloc(L * sizeof(uint16_t));
  orc_rm_table(cb_idx, rv, T);
  for (uint32_t i = 0; i < E; i++) {
    uint16_t t = T[i % L];
    buf[t]     = add16(buf[t], e[i]);
  }
  free(T);
  return 0;
}

int orc_rm_tx(const uint8_t* coded, uint32_t K, uint8_t* e, uint32_t E, uint32_t rv)
{
  int idx = orc_cbindex_exact(K);
  if (idx < 0 || rv >= 4) {
    return -2;
  }
  uint32_t  L = 3 * K + 12;
  uint16_t* T = malloc(L * sizeof(uint16_t));
  orc_rm_table((uint32_t)idx, rv, T);
  for (uint32_t i = 0; i < E; i++) {
    e[i] = coded[T[i % L]];
  }
  free(T);
  return 0;
}

/* ------------------------------------------------------------------ max-log-MAP constituent decoder */
/*
 * Trellis as (from-state, parity-bit) per destination state, split by information bit:
 * branch metric gamma(u,p) = u*x + p*y.
 */
static const uint8_t PRED_U0[8] = {0, 3, 4, 7, 1, 2, 5, 6};
static const uint8_t PAR_U0[8]  = {0, 1, 1, 0, 0, 1, 1, 0};
static const uint8_t PRED_U1[8] = {1, 2, 5, 6, 0, 3, 4, 7};
static const uint8_t PAR_U1[8]  = {1, 0, 0, 1, 1, 0, 0, 1};

typedef struct {
  int16_t* beta; /* 8*(K+4) */
} map_ws_t;

static void map_decode(map_ws_t* ws, uint32_t K, const int16_t* input, const int16_t* app, const int16_t* parity,
                       int16_t* output)
{
  int16_t* beta = ws->beta;
  int16_t  st[8], nx[8];
  uint32_t end = K + 3;

  /* backward recursion */
  st[0] = 0;
  for (int s = 1; s < 8; s++) {
    st[s] = NEG_INF;
  }
  memcpy(&beta[8 * end], st, sizeof(st));
  for (int k = (int)end - 1; k >= 0; k--) {
    int16_t x = input[k];
    if (app && (uint32_t)k < K) {
      x = add16(x, app[k]);
    }
    int16_t y      = parity[k];
    int16_t g[2][2] = {{0, y}, {x, add16(x, y)}}; /* g[u][p] */
    /* state s at time k goes to d0 (u=0) and d1 (u=1): invert the predecessor tables */
    for (int d = 0; d < 8; d++) {
      nx[d] = 0;
    }
    int16_t c0[8], c1[8];
    for (int d = 0; d < 8; d++) {
      int s0 = PRED_U0[d], s1 = PRED_U1[d];
      c0[s0] = PAR_U0[d] ? add16(st[d], g[0][1]) : st[d];
      c1[s1] = add16(st[d], g[1][PAR_U1[d]]);
    }
    for (int s = 0; s < 8; s++) {
      nx[s]           = max16(c0[s], c1[s]);
      beta[8 * k + s] = nx[s];
    }
    if ((k % 4) == 0 && (uint32_t)k < K) {
      for (int s = 1; s < 8; s++) {
        nx[s] = sub16(nx[s], nx[0]);
      }
      nx[0] = 0;
    }
    memcpy(st, nx, sizeof(st));
  }

  /* forward recursion + LLR */
  st[0] = 0;
  for (int s = 1; s < 8; s++) {
    st[s] = NEG_INF;
  }
  for (uint32_t k = 1; k <= K; k++) {
    int16_t x = input[k - 1];
    if (app) {
      x = add16(x, app[k - 1]);
    }
    int16_t        y  = parity[k - 1];
    int16_t        xy = add16(x, y);
    const int16_t* b  = &beta[8 * k];
    int16_t        b0[8], b1[8];
    for (int d = 0; d < 8; d++) {
      b0[d] = PAR_U0[d] ? add16(st[PRED_U0[d]], y) : st[PRED_U0[d]];
      b1[d] = add16(st[PRED_U1[d]], PAR_U1[d] ? xy : x);
    }
    int16_t m0 = add16(b0[0], b[0]), m1 = add16(b1[0], b[0]);
    for (int d = 1; d < 8; d++) {
      m0 = max16(m0, add16(b0[d], b[d]));
      m1 = max16(m1, add16(b1[d], b[d]));
    }
    for (int d = 0; d < 8; d++) {
      st[d] = max16(b0[d], b1[d]);
    }
    if ((k % 4) == 0) {
      for (int s = 1; s < 8; s++) {
        st[s] = sub16(st[s], st[0]);
      }
      st[0] = 0;
    }
    output[k - 1] = sub16(m1, m0);
  }
}

int orc_map_gen(uint32_t K, const int16_t* input, const int16_t* app, const int16_t* parity, int16_t* output)
{
  map_ws_t ws;
  ws.beta = malloc(sizeof(int16_t) * 8 * (K + 4));
  if (!ws.beta) {
    return -1;
  }
  map_decode(&ws, K, input, app, parity, output);
  free(ws.beta);
  return 0;
}

/* ------------------------------------------------------------------ turbo schedule */
typedef struct {
  uint32_t  K;
  uint32_t  n_iter;
  int16_t * syst, *par0, *par1, *app1, *app2, *ext1, *ext2;
  uint16_t *fwd, *rev;
  map_ws_t  ws;
  int16_t*  mem;
} tdec_t;

static int tdec_open(tdec_t* d, uint32_t K)
{
  memset(d, 0, sizeof(*d));
  if (orc_cbindex_exact(K) < 0) {
    return -1;
  }
  d->K        = K;
  size_t n    = K + 16;
  d->mem      = calloc(7 * n + 8 * (K + 4), sizeof(int16_t));
  d->syst     = d->mem;
  d->par0     = d->syst + n;
  d->par1     = d->par0 + n;
  d->app1     = d->par1 + n;
  d->app2     = d->app1 + n;
  d->ext1     = d->app2 + n;
  d->ext2     = d->ext1 + n;
  d->ws.beta  = d->ext2 + n;
  d->fwd      = malloc(2 * K * sizeof(uint16_t));
  d->rev      = d->fwd + K;
  orc_qpp(K, d->fwd, d->rev);
  return 0;
}

static void tdec_close(tdec_t* d)
{
  free(d->mem);
  free(d->fwd);
}

static void tdec_half_iteration(tdec_t* d, const int16_t* in)
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
      for (uint32_t i = 0; i < K; i++) {
        d->app1[i] = sub16(d->app1[i], d->ext1[i]);
      }
    }
    map_decode(&d->ws, K, d->syst, n ? d->app1 : NULL, d->par0, d->ext1);
  } else {
    if (n > 1) {
      for (uint32_t i = 0; i < K; i++) {
        d->ext1[i] = sub16(d->ext1[i], d->app1[i]);
      }
    }
    for (uint32_t i = 0; i < K; i++) {
      d->app2[d->rev[i]] = d->ext1[i];
    }
    map_decode(&d->ws, K, d->app2, NULL, d->par1, d->ext2);
    for (uint32_t i = 0; i < K; i++) {
      d->app1[d->fwd[i]] = d->ext2[i];
    }
  }
  d->n_iter++;
}

static void tdec_decide(const tdec_t* d, uint8_t* out)
{
  const int16_t* src = (d->n_iter % 2) ? d->ext1 : d->app1;
  for (uint32_t i = 0; i < d->K / 8; i++) {
    uint8_t b = 0;
    for (int j = 0; j < 8; j++) {
      b = (uint8_t)((b << 1) | (src[8 * i + j] > 0 ? 1 : 0));
    }
    out[i] = b;
  }
}

int orc_tdec_trace(uint32_t K, const int16_t* in, uint32_t nof_iter, uint8_t* out_bytes, int16_t* dump)
{
  tdec_t d;
  if (tdec_open(&d, K)) {
    return -1;
  }
  for (uint32_t it = 0; it < nof_iter; it++) {
    tdec_half_iteration(&d, in);
    tdec_decide(&d, &out_bytes[it * (K / 8)]);
    if (dump) {
      memcpy(&dump[(it * 3 + 0) * K], d.ext1, sizeof(int16_t) * K);
      memcpy(&dump[(it * 3 + 1) * K], d.ext2, sizeof(int16_t) * K);
      memcpy(&dump[(it * 3 + 2) * K], d.app1, sizeof(int16_t) * K);
    }
  }
  tdec_close(&d);
  return 0;
}

int orc_tdec_run_all(uint32_t K, const int16_t* in, uint32_t nof_iter, uint8_t* out_bytes)
{
  tdec_t d;
  if (tdec_open(&d, K)) {
    return -1;
  }
  do {
    tdec_half_iteration(&d, in);
  } while (d.n_iter < nof_iter);
  tdec_decide(&d, out_bytes);
  tdec_close(&d);
  return 0;
}

/* ------------------------------------------------------------------ batch (CPU baseline "port" leg + parity at scale) */
typedef struct {
  uint32_t       K, first, last, max_iter;
  int            early_stop;
  const int16_t* in;
  uint8_t *      out, *noi, *crc_ok;
} bjob_t;

static void* bjob_run(void* arg)
{
  bjob_t* j = (bjob_t*)arg;
  tdec_t  d;
  if (tdec_open(&d, j->K)) {
    return NULL;
  }
  uint32_t L = 3 * j->K + 12;
  for (uint32_t n = j->first; n < j->last; n++) {
    const int16_t* in  = &j->in[(size_t)n * L];
    uint8_t*       out = &j->out[(size_t)n * (j->K / 8)];
    d.n_iter           = 0;
    int      ok        = 0;
    uint32_t noi       = 0;
    do {
      tdec_half_iteration(&d, in);
      noi++;
      if (j->early_stop) {
        tdec_decide(&d, out);
        if (noi >= 2 && orc_crc_bytes(ORC_CRC24B, 24, out, (int)j->K) == 0) {
          ok = 1;
        }
      }
    } while (noi < j->max_iter && !ok);
    if (!j->early_stop) {
      tdec_decide(&d, out);
      ok = orc_crc_bytes(ORC_CRC24B, 24, out, (int)j->K) == 0;
    }
    j->noi[n]    = (uint8_t)noi;
    j->crc_ok[n] = (uint8_t)ok;
  }
  tdec_close(&d);
  return NULL;
}

double orc_tdec_batch(uint32_t K, const int16_t* in, uint32_t n, uint32_t max_iter, int early_stop, int nthreads,
                      uint8_t* out, uint8_t* noi, uint8_t* crc_ok)
{
  if (nthreads < 1) {
    nthreads = 1;
  }
  if (max_iter < 1) {
    max_iter = 1;
  }
  pthread_t*      th   = calloc((size_t)nthreads, sizeof(pthread_t));
  bjob_t*         jobs = calloc((size_t)nthreads, sizeof(bjob_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (bjob_t){K, (uint32_t)((uint64_t)n * t / nthreads), (uint32_t)((uint64_t)n * (t + 1) / nthreads),
                       max_iter, early_stop, in, out, noi, crc_ok};
    pthread_create(&th[t], NULL, bjob_run, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ transport-block loop */
int orc_decode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const int16_t* e_bits,
                  uint32_t max_iterations, int16_t* buffer_f, uint8_t* sb_data, uint8_t* cb_crc, uint8_t* tb_crc,
                  uint8_t* data, uint32_t* cb_noi, float* avg_iterations)
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
    int16_t* buf = &buffer_f[(size_t)r * ORC_SOFTBUFFER_SIZE];
    orc_rm_rx(&e_bits[rp], buf, E, Ki, rv);
    tdec_t d;
    tdec_open(&d, K);
    uint32_t noi = 0;
    int      ok  = 0;
    do {
      tdec_half_iteration(&d, buf);
      tdec_decide(&d, dst);
      noi++;
      avg += 1.0f;
      uint32_t crc = C > 1 ? orc_crc_bytes(ORC_CRC24B, 24, dst, (int)K) : orc_crc_bytes(ORC_CRC24A, 24, dst, (int)(tbs + 24));
      if (crc == 0 && noi >= 2) {
        cb_crc[r] = 1;
        ok        = 1;
      }
    } while (noi < max_iterations && !ok);
    cb_noi[r] = noi;
    tdec_close(&d);
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

/* ------------------------------------------------------------------ transport-block encode */
/*
 * encode_tb_off, lib/src/phy/phch/sch.c:240-358, at bit level: TB CRC24A over the payload (srsran_tcod_encode_lut feeds
 * every payload byte of every CB to crc_tb, turbocoder.c:217-225), segmentation with the K2 (smaller) blocks FIRST
 * (sch.c:287-293 - note the reference's decoder puts K1 first, sch.c:384-389; they agree only when C2 == 0, which is the
 * case for every standard TBS), CRC24B per CB when C > 1, turbo encode, rate matching with the E rule of sch.c:299-303
 * ("i <= C - gamma - 1" gets the floor), output packed at bit offset wp.
 */
int orc_encode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const uint8_t* data, uint8_t* e_bits)
{
  if (!data || !e_bits) {
    return -2;
  }
  uint32_t sg[12];
  if (orc_cbsegm(tbs, sg)) {
    return -1;
  }
  uint32_t F = sg[0], C = sg[1], K1 = sg[2], K2 = sg[3], C2 = sg[7];
  if (F || Qm == 0 || rv >= 4) {
    return -1;
  }
  memset(e_bits, 0, (nof_e_bits + 7) / 8);
  if (C == 0) {
    return 0;
  }
  uint32_t Gp = nof_e_bits / Qm, gamma = Gp % C;
  /* payload + TB CRC as bits */
  uint8_t* tb = malloc(tbs + 24);
  for (uint32_t i = 0; i < tbs; i++) {
    tb[i] = (data[i / 8] >> (7 - i % 8)) & 1;
  }
  uint32_t crc = orc_crc_bits(ORC_CRC24A, 24, tb, (int)tbs);
  for (int i = 0; i < 24; i++) {
    tb[tbs + i] = (crc >> (23 - i)) & 1;
  }
  uint8_t* cb    = malloc(ORC_MAX_K);
  uint8_t* coded = malloc(3 * ORC_MAX_K + 12);
  uint8_t* e     = malloc(nof_e_bits + 8 * Qm + 8);
  uint32_t rp = 0, wp = 0;
  int      ret = 0;
  for (uint32_t i = 0; i < C && !ret; i++) {
    uint32_t K    = i < C2 ? K2 : K1;
    uint32_t rlen = C > 1 ? K - 24 : K;
    uint32_t n_e  = (i <= C - gamma - 1) ? Qm * (Gp / C) : Qm * (uint32_t)ceilf((float)Gp / C);
    memcpy(cb, &tb[rp], rlen);
    if (C > 1) {
      uint32_t c = orc_crc_bits(ORC_CRC24B, 24, cb, (int)rlen);
      for (int b = 0; b < 24; b++) {
        cb[rlen + b] = (c >> (23 - b)) & 1;
      }
    }
    if (orc_tcod_encode(cb, coded, K) || orc_rm_tx(coded, K, e, n_e, rv)) {
      ret = -1;
      break;
    }
    for (uint32_t n = 0; n < n_e; n++) {
      uint32_t pos = wp + n;
      if (e[n]) {
        e_bits[pos / 8] |= (uint8_t)(0x80 >> (pos % 8));
      }
    }
    rp += rlen;
    wp += n_e;
  }
  free(tb);
  free(cb);
  free(coded);
  free(e);
  return ret;
}

/* ------------------------------------------------------------------ UL-SCH channel de-interleaver */
int orc_ulsch_deinterleave(const int16_t* q, uint32_t Qm, uint32_t H, uint32_t nsymb, int16_t* g, const uint32_t* ri_pos, uint32_t nof_ri)
{
  if (!q || !g || Qm == 0 || nsymb == 0 || H < nsymb) {
    return -2;
  }
  uint32_t n = H * Qm, rows = H / nsymb, cols = nsymb;
  uint8_t*  is_ri = calloc(n + 1, 1);
  uint32_t* lut   = calloc(n + 1, sizeof(uint32_t));
  for (uint32_t i = 0; i < nof_ri; i++) {
    if (ri_pos[i] < n) {
      is_ri[ri_pos[i]] = 1;
    }
  }
  uint32_t idx = 0;
  for (uint32_t j = 0; j < rows; j++) {
    for (uint32_t i = 0; i < cols; i++) {
      for (uint32_t k = 0; k < Qm; k++) {
        uint32_t p = j * Qm + i * rows * Qm + k;
        lut[p]     = is_ri[p] ? 0 : idx++;
      }
    }
  }
  /* positions outside the rows x cols matrix (H not a multiple of nsymb) keep lut = 0, as the reference's calloc'd /
   * reused table would - callers always pass H = rows * cols */
  for (uint32_t p = 0; p < n; p++) {
    g[lut[p]] = q[p];
  }
  free(is_ri);
  free(lut);
  return 0;
}

/* ------------------------------------------------------------------ scrambling sequence */
void orc_sequence_bits(uint32_t c_init, uint8_t* c, uint32_t len)
{
  /* x1(n+31) = x1(n+3) + x1(n), x1(0) = 1; x2(n+31) = x2(n+3) + x2(n+2) + x2(n+1) + x2(n), x2(0..30) = bits of c_init;
   * c(n) = x1(n + 1600) + x2(n + 1600), all mod 2. 31-bit windows: bit j of a register = x(n + j). */
  uint32_t x1 = 1, x2 = c_init & 0x7fffffffu;
  for (uint32_t n = 0; n < 1600 + len; n++) {
    if (n >= 1600) {
      c[n - 1600] = (uint8_t)((x1 ^ x2) & 1u);
    }
    uint32_t f1 = (x1 ^ (x1 >> 3)) & 1u;
    uint32_t f2 = (x2 ^ (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3)) & 1u;
    x1          = (x1 >> 1) | (f1 << 30);
    x2          = (x2 >> 1) | (f2 << 30);
  }
}

void orc_sequence_apply_s(const int16_t* in, int16_t* out, uint32_t len, uint32_t c_init)
{
  uint8_t* c = malloc(len ? len : 1);
  orc_sequence_bits(c_init, c, len);
  for (uint32_t i = 0; i < len; i++) {
    out[i] = c[i] ? (int16_t)(uint16_t)(0u - (uint16_t)in[i]) : in[i];
  }
  free(c);
}

/* ------------------------------------------------------------------ soft demodulation to int16 LLRs */
/*
 * srsran_demod_soft_demodulate_s (lib/src/phy/modem/demod_soft.c:871-894) as the reference's AVX2 / SSE build computes it. The
 * 16QAM and 64QAM bodies handle four symbols per SSE trip with ROUND-TO-NEAREST-EVEN conversion and a saturating pack
 * (_mm_cvtps_epi32, _mm_packs_epi32, :250-287 and :569-629) and the last nsymbols % 4 symbols in scalar code that TRUNCATES
 * (:290-298, :631-643) - the 16QAM tail also subtracts its threshold in floating point before truncating; QPSK goes through
 * srsran_vec_convert_fi (vector_simd.c:436-472): truncation everywhere, but the 16 values of an AVX2 trip are SATURATED by the pack
 * (simd.h:1866-1871) while the last (2 n) % 16 go through a plain C cast;
 * BPSK and 256QAM are scalar (double resp. float arithmetic, truncation). mod: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM
 * (srsran_mod_t, phy_common.h:285-292). symbols: nsymbols pairs (re, im). Returns -1 for an unknown modulation.
 */
#include <math.h>
static inline int32_t cvt_rne(float v)
{
  /* _mm_cvtps_epi32 in the default rounding mode; out of range / NaN -> the "integer indefinite" value */
  if (!(v > -2147483904.0f && v < 2147483648.0f)) {
    return INT32_MIN;
  }
  return (int32_t)lrintf(v); /* round to nearest even (default FP environment) */
}
static inline int16_t packs32(int32_t v)
{
  return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}
static inline int16_t abs16_wrap(int16_t v)
{
  return (int16_t)(v < 0 ? (uint16_t)(0u - (uint16_t)v) : (uint16_t)v); /* _mm_abs_epi16: |-32768| stays -32768 */
}
int orc_demod_soft_demodulate_s(int mod, const float* symbols, int16_t* llr, int nsymbols)
{
  switch (mod) {
    case 0: /* BPSK, demod_soft.c:96-101 */
      for (int i = 0; i < nsymbols; i++) {
        llr[i] = (int16_t)(-100 * (symbols[2 * i] + symbols[2 * i + 1]) * M_SQRT1_2);
      }
      return 0;
    case 1: { /* QPSK, :115-118 */
      const float scale = (float)(-100 * M_SQRT2);
      const int   len = 2 * nsymbols, nsimd = len - len % 16;
      for (int i = 0; i < len; i++) {
        /* SIMD body: _mm256_cvttps_epi32 (truncation) + saturating pack (simd.h:1866-1871); tail: plain C cast */
        const float v = symbols[i] * scale;
        llr[i]        = i < nsimd ? packs32((v > -2147483904.0f && v < 2147483648.0f) ? (int32_t)v : INT32_MIN) : (int16_t)v;
      }
      return 0;
    }
    case 2: { /* 16QAM, :250-299 */
      const int16_t offset = (int16_t)(2 * 400 / sqrtf(10));
      const int     nsimd  = 4 * (nsymbols / 4);
      for (int i = 0; i < nsymbols; i++) {
        for (int c = 0; c < 2; c++) {
          const float v = symbols[2 * i + c];
          if (i < nsimd) {
            const int16_t y    = packs32(cvt_rne(v * -400.0f));
            llr[4 * i + c]     = y;
            llr[4 * i + 2 + c] = (int16_t)(uint16_t)((uint16_t)abs16_wrap(y) - (uint16_t)offset);
          } else {
            const short y      = (short)(400 * v);
            llr[4 * i + c]     = (short)-y;
            llr[4 * i + 2 + c] = (short)(abs(y) - 2 * 400 / sqrtf(10));
          }
        }
      }
      return 0;
    }
    case 3: { /* 64QAM, :569-643 */
      const int16_t off1 = (int16_t)(4 * 700 / sqrtf(42)), off2 = (int16_t)(2 * 700 / sqrtf(42));
      const int     nsimd = 4 * (nsymbols / 4);
      for (int i = 0; i < nsymbols; i++) {
        for (int c = 0; c < 2; c++) {
          const float v = symbols[2 * i + c];
          if (i < nsimd) {
            const int16_t y    = packs32(cvt_rne(v * -700.0f));
            const int16_t a1   = (int16_t)(uint16_t)((uint16_t)abs16_wrap(y) - (uint16_t)off1);
            const int16_t a2   = (int16_t)(uint16_t)((uint16_t)abs16_wrap(a1) - (uint16_t)off2);
            llr[6 * i + c]     = y;
            llr[6 * i + 2 + c] = a1;
            llr[6 * i + 4 + c] = a2;
          } else {
            const int16_t y    = (int16_t)(700 * v);
            llr[6 * i + c]     = (int16_t)-y;
            llr[6 * i + 2 + c] = (int16_t)((int16_t)abs(y) - off1);
            llr[6 * i + 4 + c] = (int16_t)((int16_t)abs(llr[6 * i + 2 + c]) - off2);
          }
        }
      }
      return 0;
    }
    case 4: /* 256QAM, :824-844: float arithmetic, truncation */
      for (int i = 0; i < nsymbols; i++) {
        for (int c = 0; c < 2; c++) {
          volatile float r   = -symbols[2 * i + c];
          llr[8 * i + c]     = (int16_t)(1000 * r);
          r                  = fabsf(r) - 8.0f / sqrtf(170.0f);
          llr[8 * i + 2 + c] = (int16_t)(1000 * r);
          r                  = fabsf(r) - 4.0f / sqrtf(170.0f);
          llr[8 * i + 4 + c] = (int16_t)(1000 * r);
          r                  = fabsf(r) - 2.0f / sqrtf(170.0f);
          llr[8 * i + 6 + c] = (int16_t)(1000 * r);
        }
      }
      return 0;
    default:
      return -1;
  }
}
