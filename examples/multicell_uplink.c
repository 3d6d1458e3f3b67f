/*
 * multicell_uplink.c - the eNB uplink load of BASELINE config 5 driven from plain C through the C ABI only:
 * T worker threads (one engine each, as srsRAN runs one srsran_sch_t per PHY worker), every subframe each worker decodes
 * `cells` 100-PRB PUSCH transport blocks (TBS 75376 = 13 code blocks of K = 5824) in ONE batched submission with device-resident HARQ soft buffers. Test vectors come from the engine's own encoder
 * (srsb200_encode_tb_batch) plus AWGN; every decoded transport block is compared with its payload.
 *
 *   gcc -O2 -std=gnu99 examples/multicell_uplink.c -Iinclude -Lsrsran_4g_b200 -lsrsran_b200 -lpthread -lm \
 *       -Wl,-rpath,'$ORIGIN/../srsran_4g_b200' -o examples/multicell_uplink
 *   examples/multicell_uplink [threads per GPU=4] [cells per worker=64] [subframes=20] [pinned=1] [gpus=1]
 *
 * gpus > 1 (BASELINE config 5: 8 GPUs x 64 cells): the worker threads are spread over the devices, worker i on device i mod gpus
 * - cells, their HARQ soft-buffer mirrors and their host threads belong to one GPU for good; nothing crosses between devices
 * (independent units, no collective). Weak scaling: the work per GPU is fixed, the aggregate is reported.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "srsran_b200.h"

#define TBS 75376u
#define QM 6u
#define G_BITS 86400u
#define SOFTBUFFER_SIZE 18600
#define MAX_CB 13

static int               pinned = 1, gpus = 1;
static pthread_barrier_t g_start; /* every worker finishes its warm-up (allocations, table builds) before any is timed */

typedef struct {
  int      id, cells, subframes;
  double   seconds;
  uint64_t tb_ok, tb_total, bit_errors;
} worker_t;

static uint64_t rng_next(uint64_t* s)
{
  *s ^= *s << 13;
  *s ^= *s >> 7;
  *s ^= *s << 17;
  return *s;
}
static double gauss(uint64_t* s)
{
  double u1 = ((rng_next(s) >> 11) + 1.0) / 9007199254740993.0, u2 = (rng_next(s) >> 11) / 9007199254740992.0;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static double now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void* worker(void* arg)
{
  worker_t*         w = (worker_t*)arg;
  srsb200_engine_t* e = NULL;
  if (srsb200_engine_create(&e, w->id % gpus) != SRSB200_SUCCESS) {
    fprintf(stderr, "worker %d: %s\n", w->id, srsb200_last_error());
    return NULL;
  }
  srsb200_softbuffer_set_resident(e, 1);
  const int n       = w->cells;
  uint64_t  seed    = 0x9E3779B97F4A7C15ull * (uint64_t)(w->id + 1);
  uint8_t*  payload = malloc((size_t)n * (TBS / 8));
  uint8_t*  packed  = calloc((size_t)n, (G_BITS + 7) / 8);
  /* the two buffers that cross PCIe every subframe are page-locked (pageable ones work too, through a staging copy) */
  int16_t*  llr     = pinned ? srsb200_host_alloc((size_t)n * G_BITS * sizeof(int16_t)) : malloc((size_t)n * G_BITS * sizeof(int16_t));
  uint8_t*  out     = pinned ? srsb200_host_alloc((size_t)n * (MAX_CB * 768 + 8)) : calloc((size_t)n, MAX_CB * 768 + 8);
  /* one srsran_softbuffer_rx_t worth of host arrays per cell (the device mirror is keyed by the buffer_f pointers) */
  int16_t*  bf      = calloc((size_t)n * MAX_CB, SOFTBUFFER_SIZE * sizeof(int16_t));
  uint8_t*  sbd     = calloc((size_t)n * MAX_CB, SOFTBUFFER_SIZE / 8);
  int16_t** bfp     = malloc((size_t)n * MAX_CB * sizeof(int16_t*));
  uint8_t** sbp     = malloc((size_t)n * MAX_CB * sizeof(uint8_t*));
  uint8_t*  cb_crc  = calloc((size_t)n, MAX_CB);
  uint8_t*  tb_crc  = calloc((size_t)n, 1);
  for (int i = 0; i < n * MAX_CB; i++) {
    bfp[i] = bf + (size_t)i * SOFTBUFFER_SIZE;
    sbp[i] = sbd + (size_t)i * (SOFTBUFFER_SIZE / 8);
  }
  for (size_t i = 0; i < (size_t)n * (TBS / 8); i++) payload[i] = (uint8_t)rng_next(&seed);

  /* transmit side on the device: CRC attach, turbo encode, rate matching */
  srsb200_tb_tx_t* tx = calloc((size_t)n, sizeof(*tx));
  for (int c = 0; c < n; c++) {
    tx[c].tbs = TBS; tx[c].Qm = QM; tx[c].rv = 0; tx[c].nof_e_bits = G_BITS; tx[c].max_cb = MAX_CB;
    tx[c].data   = payload + (size_t)c * (TBS / 8);
    tx[c].e_bits = packed + (size_t)c * ((G_BITS + 7) / 8);
  }
  if (srsb200_encode_tb_batch(e, tx, (uint32_t)n) != SRSB200_SUCCESS) {
    fprintf(stderr, "worker %d: encode failed: %s\n", w->id, srsb200_last_error());
    return NULL;
  }
  /* 64QAM-like LLRs: +-700 with AWGN (rate 0.87 needs a clean channel; sigma chosen so that 5-7 half-iterations are needed) */
  for (int c = 0; c < n; c++) {
    const uint8_t* p = packed + (size_t)c * ((G_BITS + 7) / 8);
    int16_t*       l = llr + (size_t)c * G_BITS;
    for (uint32_t i = 0; i < G_BITS; i++) {
      double v = ((p[i / 8] >> (7 - i % 8)) & 1 ? 1.0 : -1.0) + 0.42 * gauss(&seed);
      l[i]     = (int16_t)(700.0 * v);
    }
  }
  srsb200_tb_t* rx = calloc((size_t)n, sizeof(*rx));
  for (int c = 0; c < n; c++) {
    rx[c].tbs = TBS; rx[c].Qm = QM; rx[c].rv = 0; rx[c].nof_e_bits = G_BITS; rx[c].max_cb = MAX_CB;
    rx[c].e_bits   = llr + (size_t)c * G_BITS;
    rx[c].buffer_f = bfp + (size_t)c * MAX_CB;
    rx[c].sb_data  = sbp + (size_t)c * MAX_CB;
    rx[c].cb_crc   = cb_crc + (size_t)c * MAX_CB;
    rx[c].tb_crc   = tb_crc + c;
    rx[c].data     = out + (size_t)c * (MAX_CB * 768 + 8);
  }
  double t0 = 0;
  for (int sf = -2; sf < w->subframes; sf++) { /* two warm-up subframes */
    if (sf == 0) {
      pthread_barrier_wait(&g_start);
      t0 = now_s();
    }
    /* new data in every cell: srsran_softbuffer_rx_reset forwarded to the device mirrors, all cells in one call */
    memset(cb_crc, 0, (size_t)n * MAX_CB);
    srsb200_softbuffer_reset(e, bfp, (uint32_t)(n * MAX_CB));
    if (srsb200_decode_tb_batch(e, rx, (uint32_t)n, 8) != SRSB200_SUCCESS) {
      fprintf(stderr, "worker %d: decode failed: %s\n", w->id, srsb200_last_error());
      return NULL;
    }
    if (sf >= 0) {
      for (int c = 0; c < n; c++) {
        w->tb_total++;
        if (rx[c].ret == 0) {
          w->tb_ok++;
          if (memcmp(rx[c].data, payload + (size_t)c * (TBS / 8), TBS / 8)) w->bit_errors++; /* CRC passed but bytes differ */
        }
      }
    }
  }
  w->seconds = now_s() - t0;
  srsb200_engine_destroy(e);
  return NULL;
}

int main(int argc, char** argv)
{
  int T = argc > 1 ? atoi(argv[1]) : 4, cells = argc > 2 ? atoi(argv[2]) : 64, sfs = argc > 3 ? atoi(argv[3]) : 20;
  if (argc > 4) pinned = atoi(argv[4]);
  if (argc > 5) gpus = atoi(argv[5]);
  if (gpus < 1 || gpus > srsb200_device_count()) {
    fprintf(stderr, "gpus=%d but %d CUDA device(s) visible\n", gpus, srsb200_device_count());
    return 2;
  }
  T *= gpus; /* threads per GPU x GPUs */
  if (T < 1 || T > 64 || cells < 1 || sfs < 1) return 2;
  pthread_t th[64];
  worker_t  w[64];
  memset(w, 0, sizeof(w));
  pthread_barrier_init(&g_start, NULL, (unsigned)T);
  for (int i = 0; i < T; i++) {
    w[i].id = i; w[i].cells = cells; w[i].subframes = sfs;
    pthread_create(&th[i], NULL, worker, &w[i]);
  }
  uint64_t ok = 0, tot = 0, bad = 0;
  double   slowest = 0;
  for (int i = 0; i < T; i++) {
    pthread_join(th[i], NULL);
    ok += w[i].tb_ok; tot += w[i].tb_total; bad += w[i].bit_errors;
    if (w[i].seconds > slowest) slowest = w[i].seconds;
  }
  if (tot == 0 || slowest <= 0) {
    fprintf(stderr, "no subframe decoded\n");
    return 1;
  }
  printf("{\"gpus\": %d, \"workload\": \"%d worker threads x %d subframes x %d cells x TBS %u (13 code blocks), 64QAM, rate 0.87, device-resident soft buffers, %s e-bit/data buffers\", "
         "\"transport_blocks\": %llu, \"crc_ok\": %llu, \"crc_ok_but_payload_differs\": %llu, \"seconds_slowest_worker\": %.6f, "
         "\"ms_per_subframe_aggregate\": %.4f, \"info_Gbit_s\": %.3f}\n",
         gpus, T, sfs, cells, TBS, pinned ? "page-locked" : "pageable", (unsigned long long)tot, (unsigned long long)ok, (unsigned long long)bad, slowest,
         slowest / ((double)T * sfs) * 1e3, (double)tot * TBS / slowest / 1e9);
  return bad == 0 ? 0 : 1;
}
