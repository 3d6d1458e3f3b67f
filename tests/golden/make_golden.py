#!/usr/bin/env python
"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libsrsran_ref.so, built from /root/reference by
oracle/Makefile) and from the known-answer vectors in the reference's own test headers. Run in the build container only:

    python tests/golden/make_golden.py

The fixtures travel to the GPU box, where /root/reference does not exist; tests/test_oracle_golden.py checks the
clean-room oracle against them and tests/test_gpu_parity.py checks the CUDA engine against them."""
import ctypes
import os
import re
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import vecgen  # noqa: E402

REF = os.environ.get("SRSRAN_REF", "/root/reference")


def parse_c_array(text, name):
    m = re.search(name + r"\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    return np.array([int(x) for x in re.findall(r"\d+", m.group(1))], np.uint8)


def main():
    ol.build_ref()
    r = ol.ref()
    assert r is not None, "needs the compiled reference"

    # ---- known-answer vectors from the reference's own tests
    hdr = open(os.path.join(REF, "lib/src/phy/fec/turbo/test/turbodecoder_test.h")).read()
    known_data = parse_c_array(hdr, "known_data")
    known_enc = parse_c_array(hdr, "known_data_encoded")
    assert len(known_data) == 504 and len(known_enc) == 3 * 504 + 12
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    crc_bits = np.array([libc.rand() % 2 for _ in range(5001)], np.uint8)  # crc_test.c:89-95, seed 1
    np.savez_compressed(os.path.join(HERE, "kat.npz"), known_data=known_data, known_data_encoded=known_enc,
                        crc_test_bits=crc_bits, crc24a=np.uint32(0x1C5C97), crc24b=np.uint32(0x36D1F0))

    # ---- tables: CRC32 fingerprints of every QPP permutation and every rate-matching table of the reference
    qpp_fp = np.zeros(188, np.uint32)
    rm_fp = np.zeros((188, 4), np.uint32)
    sizes = np.zeros(188, np.uint32)
    for idx in range(188):
        K = r.cbsize(idx)
        sizes[idx] = K
        f, rv_ = r.qpp(K)
        qpp_fp[idx] = zlib.crc32(f.tobytes() + rv_.tobytes())
        for rv in range(4):
            rm_fp[idx, rv] = zlib.crc32(r.rm_table(idx, rv).tobytes())
    full = {"rm_%d_%d" % (idx, rv): r.rm_table(idx, rv) for idx, rv in ((0, 0), (0, 3), (58, 1), (187, 0), (187, 2))}
    segs = {}
    tbs_list = [16, 40, 1000, 6120, 6200, 12216, 36696, 75376, 149776, 97896, 299856]
    seg_arr = np.array([[r.cbsegm(t)[1][k] for k in ("F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2")] for t in tbs_list], np.uint32)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), sizes=sizes, qpp_fp=qpp_fp, rm_fp=rm_fp, tbs_list=np.array(tbs_list, np.uint32),
                        seg=seg_arr, **full)

    # ---- decoder traces (generic int16, natural layout), hard bits after every half-iteration + soft fingerprints
    cases = [(40, 3.0, 100), (64, 1.0, 100), (504, 2.0, 100), (512, 0.5, 100), (1024, 1.5, 100), (2048, 1.2, 100), (3136, 1.0, 100),
             (6144, 1.5, 100), (6144, 1.0, 100), (6144, 6.0, 400), (5824, 4.0, 700), (1024, 0.0, 1000), (256, 9.0, 4000)]
    dec = {}
    for n, (K, eb, scale) in enumerate(cases):
        _, llr = vecgen.make_cb(K, eb, 5000 + n, scale)
        hard, dump = r.tdec_trace(K, llr, 10, dump=True)
        soft_fp = np.array([[zlib.crc32(dump[it, a].tobytes()) for a in range(3)] for it in range(10)], np.uint32)
        dec["llr_%d" % n] = llr
        dec["hard_%d" % n] = hard
        dec["softfp_%d" % n] = soft_fp
        dec["ext1_it3_%d" % n] = dump[3, 0]  # one full soft array per case for debugging a mismatch
    dec["cases"] = np.array(cases, np.float64)
    np.savez_compressed(os.path.join(HERE, "decoder.npz"), **dec)

    # ---- batch with CRC early stop: iteration counts + verdicts
    K = 1024
    _, llr = vecgen.make_cb_batch(K, 48, 1.2, 31)
    _, out, noi, ok = r.tdec_batch(K, llr, 8, True, nthreads=4, impl=1)
    np.savez_compressed(os.path.join(HERE, "batch_k1024.npz"), llr=llr, out=out, noi=noi, ok=ok, max_iter=np.uint32(8))

    # ---- transport blocks with HARQ retransmissions (rate de-matching + soft combining + per-CB CRC + TB CRC)
    tb = {}
    tb_cases = [(40, 300, 2, 0.5), (6200, 9000, 4, 0.5), (12216, 19200, 6, 0.5), (75376, 86400, 6, 2.0)]
    for n, (tbs, G, Qm, eb) in enumerate(tb_cases):
        st = None
        for tx, rv in enumerate((0, 2, 3, 1)):
            _, e = vecgen.make_tb(tbs, G, Qm, rv, eb, 900 + n, scale=100)
            res = r.decode_tb(tbs, Qm, rv, e, 6, st)
            st = res["state"]
            p = "tb%d_tx%d_" % (n, tx)
            tb[p + "e"] = e
            tb[p + "ret"] = np.int32(res["ret"])
            tb[p + "data"] = res["data"]
            tb[p + "noi"] = res["cb_noi"]
            tb[p + "cb_crc"] = st["cb_crc"].copy()
            tb[p + "buf_fp"] = np.uint32(zlib.crc32(st["buffer_f"].tobytes()))
            tb[p + "avg"] = np.float32(res["avg_iterations"])
    tb["cases"] = np.array(tb_cases, np.float64)
    np.savez_compressed(os.path.join(HERE, "tb_harq.npz"), **tb)
    # ---- transmit side + UL-SCH de-interleaver, from the LITERAL sch.c (srsran_dlsch_encode2, ulsch_deinterleave)
    tx = {}
    enc_cases = [(40, 2, 0, 120), (6120, 4, 1, 4 * 2500), (12216, 6, 2, 19200), (36696, 2, 3, 2 * 30000), (75376, 6, 0, 86400)]
    for n, (tbs, Qm, rv, Gb) in enumerate(enc_cases):
        data = np.random.default_rng(4000 + n).integers(0, 256, tbs // 8, dtype=np.uint8)
        ret, e = r.dlsch_encode(tbs, Qm, rv, Gb, data)
        assert ret == 0
        tx["enc%d_data" % n] = data
        tx["enc%d_e" % n] = e
    tx["enc_cases"] = np.array(enc_cases, np.uint32)
    dei_cases = [(2, 12 * 12 * 6, 12, 0), (4, 12 * 12 * 25, 12, 8), (6, 12 * 11 * 100, 11, 24)]
    for n, (Qm, H, nsymb, nri) in enumerate(dei_cases):
        rows = H // nsymb
        q = np.random.default_rng(5000 + n).integers(-30000, 30000, H * Qm).astype(np.int16)
        ri = []
        for m in range(nri // Qm if Qm else 0):
            rr, cc = rows - 1 - m // 4, (1, 4, 7, 10)[m % 4]
            ri += [rr * Qm + cc * rows * Qm + k for k in range(Qm)]
        tx["dei%d_q" % n] = q
        tx["dei%d_ri" % n] = np.array(ri, np.uint32)
        tx["dei%d_g" % n] = r.ulsch_deinterleave(q, Qm, H, nsymb, ri)
    tx["dei_cases"] = np.array(dei_cases, np.uint32)
    np.savez_compressed(os.path.join(HERE, "tx.npz"), **tx)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
