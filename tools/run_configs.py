#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not the bench line (1, 3, 4, 5) on one B200 and checks a sample of each
against the oracle. Writes profiles/r02_configs.json. Usage: python tools/run_configs.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import vecgen  # noqa: E402
import srsran_4g_b200 as sb  # noqa: E402
from srsran_4g_b200 import synth  # noqa: E402

o = ol.oracle()
eng = sb.Engine(0)
res = {}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


# ---- config 1: single code block K=6144, 4 and 8 half-iterations, no early stop (turbodecoder_test shape)
K = 6144
_, llr = vecgen.make_cb(K, 1.5, 1)
t = sb.Tdec(eng, K)
for nit in (4, 8):
    ret, out = t.run_all(llr, nit, K)
    assert ret == 0 and (out == o.tdec_run_all(K, llr, nit)).all()
    dt = timeit(lambda: t.run_all(llr, nit, K), reps=10)
    res["config1_single_cb_K6144_%d_half_iterations" % nit] = {"latency_us": dt * 1e6, "info_Mbit_s": K / dt / 1e6, "parity": "bit-exact vs oracle"}
t.free()

# ---- config 1 in the reference's 8-bit LLR mode: the windowed decoder has no K-long dependency chain (one warp, 32 windows)
l8 = np.clip(vecgen.make_cb(K, 1.5, 1, scale=12)[1], -127, 127).astype(np.int8)
for nit in (4, 8):
    out8, _, _ = eng.tdec_batch8(K, l8[None, :], nit, early_stop=False, crc_kind=sb.CRC_NONE)
    assert (out8[0] == o.tdec8_trace(K, l8, nit)[nit - 1]).all()
    dt = timeit(lambda: eng.tdec_batch8(K, l8[None, :], nit, early_stop=False, crc_kind=sb.CRC_NONE), reps=10)
    res["config1_single_cb_K6144_%d_half_iterations_llr8" % nit] = {"latency_us": dt * 1e6, "info_Mbit_s": K / dt / 1e6, "parity": "bit-exact vs the 8-bit oracle"}

# ---- config 4: all 188 LTE sizes in one submission (8 blocks each), CRC early stop
Ks, llrs = [], []
for idx in range(188):
    k = o.cbsize(idx)
    _, l = synth.make_llr_batch(k, 8, 2.0 if k < 512 else 1.5, 100 + idx, n_distinct=8)
    for i in range(8):
        Ks.append(k); llrs.append(l[i])
Ks = np.array(Ks, np.uint32)
outs, noi, ok = eng.tdec_batch(Ks, llrs, 8, early_stop=True)
for i in range(0, len(Ks), 37):
    _, oo, on, ook = o.tdec_batch(int(Ks[i]), llrs[i][None, :], 8, True)
    assert on[0] == noi[i] and ook[0] == ok[i] and (oo[0] == outs[i]).all()
dt = timeit(lambda: eng.tdec_batch(Ks, llrs, 8, early_stop=True), reps=3, warm=1)
# the same submission through the C ABI with flat, caller-prepared arrays (what a C caller pays), 8 and 64 blocks per size
import ctypes as C
Lc = sb.lib()
for per in (8, 64):
    Kf, lf = [], []
    for idx in range(188):
        k = o.cbsize(idx)
        _, l = synth.make_llr_batch(k, per, 2.0 if k < 512 else 1.5, 100 + idx, n_distinct=min(per, 8))
        for i in range(per):
            Kf.append(k); lf.append(l[i])
    Kf = np.array(Kf, np.uint32)
    flat = np.concatenate(lf).astype(np.int16)
    loff = np.concatenate([[0], np.cumsum(3 * Kf.astype(np.uint64) + 12)[:-1]]).astype(np.uint64)
    ooff = np.concatenate([[0], np.cumsum(Kf.astype(np.uint64) // 8)[:-1]]).astype(np.uint64)
    fo = np.zeros(int((Kf // 8).sum()), np.uint8); fn = np.zeros(len(Kf), np.uint8); fk = np.zeros(len(Kf), np.uint8)
    kinds = np.full(len(Kf), sb.CRC_24B, np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    def call():
        assert Lc.srsb200_tdec_batch(eng.handle, len(Kf), vp(Kf), vp(kinds), vp(flat), vp(loff), len(flat), 8, 2, 1, vp(fo), vp(ooff), len(fo), vp(fn), vp(fk)) == 0
    dtf = timeit(call, reps=5, warm=2)
    eng.profile(True); eng.profile_read(); call(); prof = eng.profile_read(); eng.profile(False)
    res["config4_mixed_188_sizes_x%d_c_abi" % per] = {"code_blocks": int(len(Kf)), "info_bits": int(Kf.sum()), "ms_host_to_host": dtf * 1e3,
                                                      "info_Mbit_s_host_to_host": float(Kf.sum()) / dtf / 1e6,
                                                      "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
                                                      "info_Mbit_s_kernels": float(Kf.sum()) / (sum(v[0] for v in prof.values()) * 1e-3) / 1e6,
                                                      "note": "pageable numpy buffers (the H2D of the LLRs dominates the 64-per-size case)"}

res["config4_mixed_188_sizes_x8"] = {"code_blocks": int(len(Ks)), "info_bits": int(Ks.sum()), "ms_host_to_host": dt * 1e3, "info_Mbit_s_host_to_host": float(Ks.sum()) / dt / 1e6,
                                     "crc_ok_fraction": float(ok.mean()), "parity": "sampled blocks bit-exact vs oracle",
                                     "note": "includes python-side concatenation of the ragged inputs and plan building for 188 buckets"}

# ---- config 3: PDSCH 100 PRB 64QAM MCS 28: TBS 75376 per codeword (C=13, K=5824), G=86400, 2 codewords, HARQ rv 0 then 2
tbs, G, Qm = 75376, 86400, 6
reqs_tx = []
for cw in range(2):
    tb = sb.TransportBlock(tbs)
    es = [vecgen.make_tb(tbs, G, Qm, rv, 4.0, 50 + cw, scale=700)[1] for rv in (0, 2)]
    reqs_tx.append((tb, es))
st = [None, None]
for tx, rv in enumerate((0, 2)):
    reqs = [(tb, Qm, rv, es[tx]) for tb, es in reqs_tx]
    t0 = time.perf_counter()
    assert eng.decode_tb_batch(reqs, 8) == 0
    dt = time.perf_counter() - t0
    for cw, (tb, es) in enumerate(reqs_tx):
        r = o.decode_tb(tbs, Qm, rv, es[tx], 8, st[cw]); st[cw] = r["state"]
        assert tb.ret == r["ret"] and (tb.cb_noi[:13] == r["cb_noi"][:13]).all() and (tb.buffer_f == st[cw]["buffer_f"]).all()
    res["config3_pdsch_2cw_tx%d_rv%d" % (tx, rv)] = {"ret": [tb.ret for tb, _ in reqs_tx], "ms_first_call": dt * 1e3, "avg_half_iterations": [tb.avg_iterations for tb, _ in reqs_tx],
                                                    "parity": "return code, iteration counts, soft buffers bit-exact vs oracle"}
# steady-state timing of the 2-codeword decode (fresh soft buffers every call)
def one_pdsch():
    rq = []
    for cw in range(2):
        tb = reqs_tx[cw][0]
        tb.buffer_f[:] = 0; tb.cb_crc[:] = 0
        rq.append((tb, Qm, 0, reqs_tx[cw][1][0]))
    eng.decode_tb_batch(rq, 8)
dt = timeit(one_pdsch, reps=10)
res["config3_pdsch_2cw_steady"] = {"ms_per_subframe_host_to_host": dt * 1e3, "info_Mbit_s": 2 * tbs / dt / 1e6, "code_blocks": 26,
                                   "h2d_bytes_per_info_bit": 2 * (G * 2 + 16 * 18600 * 2) / float(2 * tbs),
                                   "note": "host-coherent soft buffers: 16 x 18600 int16 per codeword cross PCIe both ways on top of the e-bits"}

# ---- config 3 from equalised SYMBOLS: pdsch.c:693-740 - soft demodulation, descrambling, rate de-matching, decode - in
# one submission; the LLRs never exist on the host. 8 bytes per resource element in instead of 2 bytes per LLR (12 per 64QAM RE).
import ctypes as C
from srsran_4g_b200.binding import _TbStruct
def modulate64(bits):
    b = 1.0 - 2.0 * np.asarray(bits, np.float64).reshape(-1, 6)
    ax = lambda c0: b[:, c0] * (4.0 - b[:, c0 + 2] * (2.0 - b[:, c0 + 4]))
    return ((ax(0) + 1j * ax(1)) / np.sqrt(42.0)).astype(np.complex64)
nre = G // Qm
syms, exp_sym = [], []
for cw in range(2):
    c_init = (0x1234 << 14) | (cw << 13) | (3 << 9) | 77
    payload, e_clean = vecgen.make_tb(tbs, G, Qm, 0, 60.0, 300 + cw, scale=8)
    scr = o.sequence_apply_s(np.ones(G, np.int16), c_init) < 0
    sy = modulate64((e_clean > 0).astype(np.uint8) ^ scr.astype(np.uint8))
    rngs = np.random.default_rng(900 + cw)
    sg = vecgen.sigma_for(12.5, tbs / float(G)) / np.sqrt(float(Qm))
    sy = (sy + sg * (rngs.standard_normal(nre) + 1j * rngs.standard_normal(nre))).astype(np.complex64)
    syms.append((sy, c_init))
    exp_sym.append(o.decode_tb(tbs, Qm, 0, o.sequence_apply_s(o.demod_soft_demodulate_s(3, sy), c_init), 8))
eng.softbuffer_set_resident(True)
tbs_sym = [sb.TransportBlock(tbs) for _ in range(2)]
arr = (_TbStruct * 2)()
for cw in range(2):
    tbs_sym[cw].fill_symbols(arr[cw], Qm, 0, syms[cw][0], 3, G, c_init=syms[cw][1])
def one_pdsch_symbols():
    for tb in tbs_sym:
        eng.softbuffer_reset(tb)
    assert sb.lib().srsb200_decode_tb_batch(eng.handle, arr, 2, 8) == 0
one_pdsch_symbols()
for cw in range(2):
    assert arr[cw].ret == exp_sym[cw]["ret"] == 0 and (tbs_sym[cw].cb_noi[:13] == exp_sym[cw]["cb_noi"][:13]).all()
    assert (tbs_sym[cw].data[:tbs // 8] == exp_sym[cw]["data"][:tbs // 8]).all()
dt = timeit(one_pdsch_symbols, reps=10)
res["config3_pdsch_2cw_from_symbols_steady"] = {
    "ms_per_subframe_host_to_host": dt * 1e3, "info_Mbit_s": 2 * tbs / dt / 1e6, "code_blocks": 26,
    "h2d_bytes_per_info_bit": 2 * nre * 8 / float(2 * tbs), "h2d_bytes_per_info_bit_llr_path": 2 * G * 2 / float(2 * tbs),
    "parity": "return codes, iteration counts, payload bit-exact vs the oracle chain demodulate -> descramble -> decode_tb",
    "note": "equalised symbols in (cf_t, 8 B per RE), soft buffers resident on the device: symbols -> LLRs -> descrambling -> rate de-matching -> decode without the LLRs ever existing on the host"}
eng.softbuffer_set_resident(False)

# ---- config 3, the two-layer single-TB variant (TBS 149776, C = 25, K = 6016, Qm * Nl = 12) and the 8-bit mode of both
tbs2, G2 = 149776, 12 * 14400
tb2 = sb.TransportBlock(tbs2)
e2 = vecgen.make_tb(tbs2, G2, 12, 0, 5.5, 61, scale=700)[1]
r2 = o.decode_tb(tbs2, 12, 0, e2, 8)
assert eng.decode_tb(tb2, 12, 0, e2, 8) == r2["ret"] and (tb2.cb_noi[:25] == r2["cb_noi"][:25]).all()
def one_2layer():
    tb2.buffer_f[:] = 0; tb2.cb_crc[:] = 0
    eng.decode_tb(tb2, 12, 0, e2, 8)
dt = timeit(one_2layer, reps=10)
res["config3_pdsch_2layer_1tb_steady"] = {"ms_per_subframe_host_to_host": dt * 1e3, "info_Mbit_s": tbs2 / dt / 1e6, "code_blocks": 25, "ret": int(tb2.ret),
                                          "parity": "return code and iteration counts bit-exact vs oracle"}
for name, (tbsx, Gx, Qx, nblk) in {"2cw": (tbs, G, Qm, 26), "2layer_1tb": (tbs2, G2, 12, 25)}.items():
    ntb = 2 if name == "2cw" else 1
    tbx = [sb.TransportBlock(tbsx) for _ in range(ntb)]
    ex = [np.clip(vecgen.make_tb(tbsx, Gx, Qx, 0, 1.0, 70 + c, scale=12)[1], -127, 127).astype(np.int8) for c in range(ntb)]
    rx = [o.decode_tb8(tbsx, Qx, 0, ex[c], 8) for c in range(ntb)]
    assert eng.decode_tb_batch([(tbx[c], Qx, 0, ex[c]) for c in range(ntb)], 8, llr8=True) == 0
    for c in range(ntb):
        assert tbx[c].ret == rx[c]["ret"] and (tbx[c].cb_noi[:rx[c]["seg"]["C"]] == rx[c]["cb_noi"][:rx[c]["seg"]["C"]]).all()
    def one8():
        for c in range(ntb):
            tbx[c].buffer_f[:] = 0; tbx[c].cb_crc[:] = 0
        eng.decode_tb_batch([(tbx[c], Qx, 0, ex[c]) for c in range(ntb)], 8, llr8=True)
    dt = timeit(one8, reps=10)
    res["config3_pdsch_%s_steady_llr8" % name] = {"ms_per_subframe_host_to_host": dt * 1e3, "info_Mbit_s": ntb * tbsx / dt / 1e6, "code_blocks": nblk,
                                                 "ret": [int(t_.ret) for t_ in tbx], "avg_half_iterations": [float(t_.avg_iterations) for t_ in tbx],
                                                 "parity": "return codes and iteration counts bit-exact vs the 8-bit oracle"}

# ---- config 5: 64 cells x one 100-PRB PUSCH TB (13 CB, K=5824) per subframe on one GPU, one batched submission per subframe
cells = 64
el = [vecgen.make_tb(tbs, G, Qm, 0, 6.0, 500 + c, scale=700)[1] for c in range(4)]
r0 = o.decode_tb(tbs, Qm, 0, el[0], 8)


def subframe_runner(engine, resident):
    tbl = [sb.TransportBlock(tbs) for _ in range(cells)]
    engine.softbuffer_set_resident(resident)

    def one_subframe():
        rq = []
        for c in range(cells):
            tb = tbl[c]
            if resident:
                engine.softbuffer_reset(tb)  # new-data indicator: srsran_softbuffer_rx_reset forwarded to the device mirror
            else:
                tb.buffer_f[:] = 0
                tb.cb_crc[:] = 0
            rq.append((tb, Qm, 0, el[c % 4]))
        engine.decode_tb_batch(rq, 8)
    return tbl, one_subframe


for resident in (False, True):
    tbl, one_subframe = subframe_runner(eng, resident)
    one_subframe()
    assert tbl[0].ret == r0["ret"] and (tbl[0].cb_noi[:13] == r0["cb_noi"][:13]).all() and (tbl[0].data[:9425] == r0["data"][:9425]).all()
    dt = timeit(one_subframe, reps=5, warm=1)
    res["config5_64_cells_832_cb_per_subframe_%s" % ("device_resident_softbuffers" if resident else "host_coherent_softbuffers")] = {
        "ms_per_subframe_host_to_host": dt * 1e3, "info_Mbit_s": cells * tbs / dt / 1e6, "code_blocks": 13 * cells,
        "tb_ok": int(sum(tb.ret == 0 for tb in tbl)), "parity": "cell 0 bit-exact vs oracle (ret, iteration counts, bytes)"}
eng.softbuffer_set_resident(False)

# several PHY worker threads, one engine each (the reference runs 3-4 workers): subframes of different workers overlap on the GPU
import threading
def worker(fn, n):
    for _ in range(n):
        fn()
for nthr in (4, 8):
    engs = [sb.Engine(0) for _ in range(nthr)]
    runners = [subframe_runner(engs[i], True)[1] for i in range(nthr)]
    for r in runners:
        r()
    nsub = 6
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(runners[i], nsub)) for i in range(nthr)]
    [t.start() for t in ths]; [t.join() for t in ths]
    dt = time.perf_counter() - t0
    res["config5_%d_worker_threads_device_resident" % nthr] = {"subframes": nthr * nsub, "ms_per_subframe_aggregate": dt / (nthr * nsub) * 1e3,
                                                       "info_Mbit_s": nthr * nsub * cells * tbs / dt / 1e6,
                                                       "note": "1/2/4/8-GPU scaling shards cells per GPU with no collective (one engine per worker thread per GPU)"}
    for e_ in engs:
        e_.close()
eng.close()
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
for d_ in ("profiles", "gpurun_out"):
    os.makedirs(os.path.join(ROOT, d_), exist_ok=True)
    with open(os.path.join(ROOT, d_, "r02_configs.json"), "w") as f:
        json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
