"""Parity tests proper: the CUDA engine, called through the C ABI, against the oracle on the same seeded inputs and
against the committed golden fixtures. Integer work => the bar is bit-exact (hard bits, half-iteration counts, CRC
verdicts, soft-buffer contents)."""
import os
import zlib

import numpy as np
import pytest

import oracle_lib as ol
import vecgen

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sb():
    import srsran_4g_b200 as sb
    return sb


@pytest.fixture(scope="module")
def eng(sb):
    e = sb.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def o():
    return ol.oracle()


def test_golden_decoder_traces_hard_bits(sb, eng):
    """hard decisions after every half-iteration 1..10 vs the compiled reference's (generic int16) fixtures"""
    d = np.load(os.path.join(G, "decoder.npz"))
    for n, (K, eb, scale) in enumerate(d["cases"]):
        K = int(K)
        llr = d["llr_%d" % n][None, :]
        for it in range(1, 11):
            out, noi, _ = eng.tdec_batch(K, llr, it, early_stop=False, crc_kind=sb.CRC_NONE)
            assert noi[0] == it
            assert (out[0] == d["hard_%d" % n][it - 1]).all(), (K, eb, scale, it)


def test_golden_batch_early_stop(sb, eng):
    b = np.load(os.path.join(G, "batch_k1024.npz"))
    out, noi, ok = eng.tdec_batch(1024, b["llr"], int(b["max_iter"]), early_stop=True)
    assert (noi == b["noi"]).all() and (ok == b["ok"]).all() and (out == b["out"]).all()


@pytest.mark.parametrize("idx", list(range(0, 188, 11)) + [186, 187])
def test_all_regimes_vs_oracle(sb, eng, o, idx):
    """normal + int16-overflow regimes: the degenerate wrap behaviour must be reproduced"""
    K = o.cbsize(idx)
    for eb, scale in ((1.5, 100), (6.0, 400), (0.0, 1000), (9.0, 4000)):
        _, llr = vecgen.make_cb_batch(K, 3, eb, 2000 + idx, scale)
        for it in (1, 2, 5, 8):
            out, _, _ = eng.tdec_batch(K, llr, it, early_stop=False, crc_kind=sb.CRC_NONE)
            for c in range(3):
                assert (out[c] == o.tdec_trace(K, llr[c], it)[it - 1]).all(), (K, eb, scale, it, c)


def test_group_packing_odd_counts(sb, eng, o):
    """1, 2, 33, 64, 65, 97 blocks: half-empty lanes, partial and multiple groups"""
    K = 256
    _, llr_all = vecgen.make_cb_batch(K, 97, 1.0, 5)
    _, out_o, noi_o, ok_o = o.tdec_batch(K, llr_all, 6, True, nthreads=4)
    for n in (1, 2, 33, 64, 65, 97):
        out, noi, ok = eng.tdec_batch(K, llr_all[:n], 6, early_stop=True)
        assert (noi == noi_o[:n]).all() and (ok == ok_o[:n]).all() and (out == out_o[:n]).all(), n


def test_mixed_sizes_one_submission(sb, eng, o):
    """config 4 shape: all 188 LTE sizes in one batch, mixed CRC kinds"""
    Ks, llrs, kinds = [], [], []
    for idx in range(188):
        K = o.cbsize(idx)
        for rep in range(2 if K < 1000 else 1):
            _, llr = vecgen.make_cb(K, 2.0 if K < 512 else 1.5, 7000 + 3 * idx + rep)
            Ks.append(K); llrs.append(llr); kinds.append(sb.CRC_24B if (idx + rep) % 3 else sb.CRC_24A)
    outs, noi, ok = eng.tdec_batch(np.array(Ks), llrs, 8, early_stop=True, crc_kind=np.array(kinds, np.uint8))
    for i, (K, llr, kind) in enumerate(zip(Ks, llrs, kinds)):
        hard = o.tdec_trace(K, llr, 8)
        poly = ol.CRC24B if kind == sb.CRC_24B else ol.CRC24A
        n_exp, ok_exp = 8, 0
        for it in range(1, 9):
            c = o.crc_bytes(poly, hard[it - 1], K)
            if it >= 2 and c == 0:
                n_exp, ok_exp = it, 1
                break
        if not ok_exp:
            ok_exp = int(o.crc_bytes(poly, hard[7], K) == 0)
        assert noi[i] == n_exp and ok[i] == ok_exp, (i, K)
        assert (outs[i] == hard[n_exp - 1]).all(), (i, K)


def test_tdec_object_api(sb, eng, o):
    """srsran_tdec_new_cb / _iteration / _run_all / _get_nof_iterations semantics (turbodecoder.c:510-549)"""
    t = sb.Tdec(eng, 6144)
    assert t.new_cb(41) == -1 and t.new_cb(6208) == -1
    for K in (40, 504, 6144):
        _, llr = vecgen.make_cb(K, 1.5, 31 + K)
        ref_hard = o.tdec_trace(K, llr, 6)
        assert t.new_cb(K) == 0 and t.get_nof_iterations() == 0
        for it in range(6):
            assert (t.iteration(llr, K) == ref_hard[it]).all(), (K, it)
            assert t.get_nof_iterations() == it + 1
        for n in (0, 1, 4):
            ret, out = t.run_all(llr, n, K)
            assert ret == 0 and (out == o.tdec_run_all(K, llr, n)).all()
    t2 = sb.Tdec(eng, 1024)
    assert t2.new_cb(2048) == -1
    t.free(); t2.free()


def test_large_batch_properties(sb, eng, o):
    """BASELINE config-2 shape at reduced count on the host path: 1024 blocks K=6144, true Eb/N0 1.5 dB. Size-independent
    checks: every block with crc_ok re-checks to CRC zero and equals the transmitted payload; a sample is compared with
    the oracle bit for bit; iteration-count histogram is in the range the reference shows."""
    K, n = 6144, 1024
    rng = np.random.default_rng(12)
    bits, llr = vecgen.make_cb_batch(K, 16, 1.5, 77)
    reps = n // 16
    # 16 distinct codewords, fresh noise per block
    coded = np.stack([o.encode(b) for b in bits])
    s = 2.0 * np.tile(coded, (reps, 1)).astype(np.float64) - 1.0
    big = vecgen.quantise(s + vecgen.sigma_for(1.5) * rng.standard_normal(s.shape), 100)
    out, noi, ok = eng.tdec_batch(K, big, 8, early_stop=True)
    assert ok.mean() > 0.97
    tx = np.packbits(np.tile(bits, (reps, 1)), axis=1)
    good = ok == 1
    assert (out[good] == tx[good]).all()
    assert noi.min() >= 2 and 3.5 < noi.mean() < 6.5
    for i in range(0, n, 97):
        _, oo, on, ook = o.tdec_batch(K, big[i:i + 1], 8, True)
        assert on[0] == noi[i] and ook[0] == ok[i] and (oo[0] == out[i]).all()


def test_rm_rx_lut(sb, eng, o):
    rng = np.random.default_rng(3)
    for idx, E in ((0, 100), (0, 132), (0, 400), (40, 1100), (187, 6646), (187, 18444), (187, 40000), (100, 7)):
        for rv in range(4):
            e = rng.integers(-30000, 30000, E).astype(np.int16)
            b0 = rng.integers(-3000, 3000, ol.SOFTBUFFER_SIZE).astype(np.int16)
            bo, bg = b0.copy(), b0.copy()
            assert o.rm_rx(e, bo, idx, rv) == 0 and eng.rm_turbo_rx_lut(e, bg, idx, rv) == 0
            assert (bo == bg).all(), (idx, E, rv)
    assert eng.rm_turbo_rx_lut(np.zeros(4, np.int16), np.zeros(ol.SOFTBUFFER_SIZE, np.int16), 188, 0) == -2
    assert eng.rm_turbo_rx_lut(np.zeros(4, np.int16), np.zeros(ol.SOFTBUFFER_SIZE, np.int16), 0, 4) == -2


# ---------------------------------------------------------------- transport blocks (decode_tb: sch.c:371-494, 509-573)
def _check_tb(res_o, tb, st):
    assert tb.ret == res_o["ret"]
    assert int(tb.tb_crc[0]) == res_o["tb_crc"]
    C = res_o["seg"]["C"]
    assert (tb.cb_noi[:C] == res_o["cb_noi"][:C]).all()
    assert np.float32(tb.avg_iterations) == np.float32(res_o["avg_iterations"])
    assert (tb.cb_crc[:C] == st["cb_crc"][:C]).all()
    K1 = res_o["seg"]["K1"]
    nbytes = (C - 1) * ((K1 - 24) // 8) + K1 // 8 if C > 1 else K1 // 8
    assert (tb.data[:nbytes] == res_o["data"][:nbytes]).all()
    for r in range(C):
        L = 3 * K1 + 12
        assert (tb.buffer_f[r, :L] == st["buffer_f"][r, :L]).all(), r
        assert (tb.sb_data[r] == st["sb_data"][r]).all(), r


@pytest.mark.parametrize("tbs,G,Qm,eb", [(40, 300, 2, 0.5), (6120, 14400, 2, 0.5), (6200, 9000, 4, 0.5), (12216, 19200, 6, 0.5),
                                         (36696, 43200, 6, 0.5), (75376, 86400, 6, 2.0)])
def test_decode_tb_harq_vs_oracle(sb, eng, o, tbs, G, Qm, eb):
    """rate de-matching + HARQ soft combining over rv 0,2,3,1 + per-CB CRC early stop + TB CRC24A, state carried across
    transmissions in the caller's soft buffer exactly as srsran_softbuffer_rx_t"""
    tb = sb.TransportBlock(tbs)
    st = None
    for tx, rv in enumerate((0, 2, 3, 1)):
        _, e = vecgen.make_tb(tbs, G, Qm, rv, eb, 77, scale=100)
        res_o = o.decode_tb(tbs, Qm, rv, e, 6, st)
        st = res_o["state"]
        tb.data[:] = 0  # the oracle harness hands a fresh zeroed `data` to every call; bytes a call does not write stay as they were
        assert eng.decode_tb(tb, Qm, rv, e, 6) == res_o["ret"]
        _check_tb(res_o, tb, st)


@pytest.mark.parametrize("tbs,G,Qe,eb", [(149776, 12 * 14400, 12, 5.5),   # BASELINE config 3: 64QAM x 2 layers, C = 25, K = 6016
                                          (75376, 8 * 11000, 8, 4.5),       # 256QAM
                                          (97896, 16 * 7200, 16, 4.0),      # 256QAM x 2 layers
                                          (36696, 8 * 5500, 8, 3.0)])       # 16QAM x 2 layers
def test_decode_tb_layers_codewords(sb, eng, o, tbs, G, Qe, eb):
    """decode_tb with Qm * Nl as srsran_dlsch_decode2 passes it for two layers / 256QAM (sch.c:587-604): rv 0,2,3,1 against the
    oracle, whose loop is pinned to the literal srsran_dlsch_decode2 for these very cases (tests/test_sch_literal.py)"""
    tb = sb.TransportBlock(tbs)
    st = None
    rets = []
    for tx, rv in enumerate((0, 2, 3, 1)):
        _, e = vecgen.make_tb(tbs, G, Qe, rv, eb, 177 + Qe, scale=100)
        res_o = o.decode_tb(tbs, Qe, rv, e, 8, st)
        st = res_o["state"]
        tb.data[:] = 0
        assert eng.decode_tb(tb, Qe, rv, e, 8) == res_o["ret"]
        rets.append(res_o["ret"])
        _check_tb(res_o, tb, st)
    assert 0 in rets
    if tb.seg["C"] > 16:
        # the stock 110-PRB soft buffer (max_cb = 16) cannot take this transport block: -2 like decode_tb (sch.c:541-545)
        small = sb.TransportBlock(tbs, max_cb=16)
        assert eng.decode_tb(small, Qe, 0, np.zeros(G, np.int16), 8) == -2


def test_decode_tb_golden_fixtures(sb, eng):
    """against outputs of the compiled reference (tests/golden/tb_harq.npz)"""
    t = np.load(os.path.join(G, "tb_harq.npz"))
    for n, (tbs, Gb, Qm, eb) in enumerate(t["cases"]):
        tb = sb.TransportBlock(int(tbs))
        for tx, rv in enumerate((0, 2, 3, 1)):
            p = "tb%d_tx%d_" % (n, tx)
            tb.data[:] = 0
            ret = eng.decode_tb(tb, int(Qm), rv, t[p + "e"], 6)
            assert ret == int(t[p + "ret"])
            C = tb.seg["C"]
            assert (tb.cb_noi[:C] == t[p + "noi"][:C]).all()
            assert (tb.cb_crc[:C] == t[p + "cb_crc"][:C]).all()
            assert zlib.crc32(tb.buffer_f.tobytes()) == int(t[p + "buf_fp"])
            assert np.float32(tb.avg_iterations) == t[p + "avg"]
            K1 = tb.seg["K1"]
            nbytes = (C - 1) * ((K1 - 24) // 8) + K1 // 8 if C > 1 else K1 // 8
            assert (tb.data[:nbytes] == t[p + "data"][:nbytes]).all()


def test_decode_tb_batch_many_tbs_one_submission(sb, eng, o):
    """config 3/5 shape: transport blocks of several UEs / cells / subframes in ONE batched submission"""
    cases = [(75376, 86400, 6, 6.0), (36696, 43200, 6, 1.0), (6200, 9000, 4, 0.2), (40, 300, 2, 3.0), (12216, 19200, 6, 2.5), (6120, 14400, 2, 3.0)]
    tbs_objs, reqs, exp = [], [], []
    for i, (tbs, Gb, Qm, eb) in enumerate(cases):
        _, e = vecgen.make_tb(tbs, Gb, Qm, 0, eb, 300 + i, scale=100 if Qm < 6 else 700)
        tb = sb.TransportBlock(tbs)
        tbs_objs.append(tb)
        reqs.append((tb, Qm, 0, e))
        exp.append(o.decode_tb(tbs, Qm, 0, e, 8))
    assert eng.decode_tb_batch(reqs, 8) == 0
    for tb, res_o in zip(tbs_objs, exp):
        _check_tb(res_o, tb, res_o["state"])
    assert any(r["ret"] == 0 for r in exp) and any(r["ret"] == -1 for r in exp)


def test_decode_tb_batch_big_enough_to_regroup(sb, o):
    """340 transport blocks of 13 code blocks of one size in one submission: 70 groups, so the plan regroups its unfinished code
    blocks mid-decode. One block in eight runs on a bad channel, the limits differ per transport block, two transmissions with the
    HARQ buffers on the device - every transport block against the oracle (return code, per-block iteration counts, flags, bytes)"""
    tbs, G, Qm, n = 75376, 86400, 6, 340
    rng = np.random.default_rng(1234)
    e = sb.Engine(0)
    try:
        e.softbuffer_set_resident(True)
        bad = rng.random(n) < 0.125
        limits = [int(x) for x in rng.choice([6, 8, 10], n)]
        vec = {}
        for rv in (0, 2):
            for b in (False, True):
                for v in range(3):
                    vec[(rv, b, v)] = vecgen.make_tb(tbs, G, Qm, rv, 3.8 if b else 6.0, 9000 + 10 * v + (1 if b else 0), scale=400)[1]
        which = rng.integers(0, 3, n)
        tbl = [sb.TransportBlock(tbs) for _ in range(n)]
        for tb in tbl:
            e.softbuffer_reset(tb)
        st = [None] * n
        memo = {}
        for rv in (0, 2):
            reqs = [(tbl[i], Qm, rv, vec[(rv, bool(bad[i]), int(which[i]))]) for i in range(n)]
            assert e.decode_tb_batch(reqs, 8, limits=limits) == 0
            if rv == 0:
                assert any(e.plan_regroup_points(None)), "the first transmission (4420 code blocks, one in eight slow) did not regroup"
            for i in range(n):
                key = (rv, bool(bad[i]), int(which[i]), limits[i], None if st[i] is None else st[i]["key"])
                if key not in memo:   # (the oracle result depends on the vector, the limit and the HARQ history only)
                    memo[key] = o.decode_tb(tbs, Qm, rv, vec[(rv, bool(bad[i]), int(which[i]))], limits[i], None if st[i] is None else st[i]["state"])
                res = memo[key]
                tb = tbl[i]
                assert tb.ret == res["ret"] and int(tb.tb_crc[0]) == res["tb_crc"], (rv, i)
                assert (tb.cb_noi[:13] == res["cb_noi"][:13]).all() and (tb.cb_crc[:13] == res["state"]["cb_crc"][:13]).all(), (rv, i)
                if res["ret"] == 0:
                    assert (tb.data[:tbs // 8] == res["data"][:tbs // 8]).all(), (rv, i)
                st[i] = {"state": res["state"], "key": key}
        rets = [tb.ret for tb in tbl]
        assert 0 in rets
    finally:
        e.close()


def test_decode_tb_argument_errors(sb, eng):
    """return-code mapping of decode_tb (sch.c:519-545): -2 invalid inputs, 0 for tbs == 0"""
    tb = sb.TransportBlock(6200)
    e = np.zeros(9000, np.int16)
    assert eng.decode_tb(tb, 0, 0, e, 6) == -2            # Qm == 0
    tb0 = sb.TransportBlock(0)
    assert eng.decode_tb(tb0, 2, 0, e, 6) == 0            # tbs == 0 -> SRSRAN_SUCCESS without decoding
    small = sb.TransportBlock(75376, max_cb=2)            # C = 13 > softbuffer->max_cb
    assert eng.decode_tb(small, 6, 0, np.zeros(86400, np.int16), 6) == -2
    assert eng.decode_tb(sb.TransportBlock(6128), 2, 0, np.zeros(14400, np.int16), 6) == -2   # filler bits (F != 0)


def test_decode_tb_device_resident_softbuffers(sb, o):
    """srsb200_softbuffer_set_resident(1): the HARQ buffers live on the device between transmissions; results and (after
    sync_to_host) buffer contents must equal the host-coherent / oracle ones, including reset and first-touch adoption"""
    eng = sb.Engine(0)
    eng.softbuffer_set_resident(True)
    for tbs, G, Qm, eb in ((6200, 9000, 4, 0.5), (36696, 43200, 6, 0.5)):
        tb = sb.TransportBlock(tbs)
        for round_ in range(2):  # second round re-uses the same soft buffer after a reset (new data indicator)
            eng.softbuffer_reset(tb)
            st = None
            for tx, rv in enumerate((0, 2, 3, 1)):
                _, e = vecgen.make_tb(tbs, G, Qm, rv, eb, 77 + round_, scale=100)
                res_o = o.decode_tb(tbs, Qm, rv, e, 6, st)
                st = res_o["state"]
                tb.data[:] = 0
                assert eng.decode_tb(tb, Qm, rv, e, 6) == res_o["ret"]
                eng.softbuffer_sync_to_host(tb)
                _check_tb(res_o, tb, st)
        eng.softbuffer_release(tb)
    # first touch without a reset adopts the host content
    tbs, G, Qm = 6200, 9000, 4
    tb = sb.TransportBlock(tbs)
    rng = np.random.default_rng(5)
    tb.buffer_f[:] = rng.integers(-50, 50, tb.buffer_f.shape).astype(np.int16)
    st = ol.new_tb_state(2)
    st["buffer_f"][:] = tb.buffer_f
    _, e = vecgen.make_tb(tbs, G, Qm, 0, 1.5, 3)
    res_o = o.decode_tb(tbs, Qm, 0, e, 6, st)
    assert eng.decode_tb(tb, Qm, 0, e, 6) == res_o["ret"]
    eng.softbuffer_sync_to_host(tb)
    _check_tb(res_o, tb, res_o["state"])
    eng.close()


# ------------------------------------------------------------------ transport-block encode
ENC_TBS = [16, 40, 104, 1000, 2984, 6120, 6200, 12216, 36696, 75376]


@pytest.mark.parametrize("Qm,rv", [(2, 0), (4, 1), (6, 2), (2, 3)])
def test_encode_tb_matches_oracle(eng, o, Qm, rv):
    """every e-bit of srsb200_encode_tb equals the oracle's (itself pinned to the reference's LUT encoder / rate matcher)"""
    for tbs in ENC_TBS:
        ret, seg = o.cbsegm(tbs)
        if ret or seg["F"]:
            continue
        rng = np.random.default_rng(tbs + 13 * Qm + rv)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        for G in (Qm * ((tbs * 2) // Qm), Qm * ((tbs * 4 + 12 * seg["C"]) // Qm) + 1, Qm * (tbs // Qm // 2 + 7)):
            r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
            r1, e1 = eng.encode_tb(tbs, Qm, rv, G, data)
            assert r0 == r1 == 0
            assert np.array_equal(e0, e1), f"tbs={tbs} G={G}: first differing byte {np.flatnonzero(e0 != e1)[:4]}"


def test_encode_tb_c2_order(eng, o):
    """a non-standard TBS with C2 > 0 and F == 0: the smaller code blocks go first (sch.c:287-293)"""
    for tbs in range(6200, 30000, 8):
        ret, seg = o.cbsegm(tbs)
        if ret == 0 and seg["F"] == 0 and seg["C2"] > 0:
            break
    data = np.random.default_rng(tbs).integers(0, 256, tbs // 8, dtype=np.uint8)
    for Qm, G in ((2, 6 * tbs), (6, 6 * (tbs // 4))):
        r0, e0 = o.encode_tb(tbs, Qm, 0, G, data)
        r1, e1 = eng.encode_tb(tbs, Qm, 0, G, data)
        assert r0 == r1 == 0 and np.array_equal(e0, e1)


def test_encode_tb_batch_mixed(eng, o):
    """many transport blocks of different shapes in one submission"""
    rng = np.random.default_rng(77)
    reqs = []
    for tbs, Qm, rv in [(75376, 6, 0), (16, 2, 1), (36696, 4, 2), (6120, 2, 3), (2984, 6, 0), (12216, 4, 0), (75376, 2, 1)] * 3:
        G = Qm * int(rng.integers(tbs // Qm // 2 + 8, 3 * tbs // Qm + 40))
        reqs.append((tbs, Qm, rv, G, rng.integers(0, 256, tbs // 8, dtype=np.uint8)))
    ret, res = eng.encode_tb_batch(reqs)
    assert ret == 0
    for (tbs, Qm, rv, G, data), (r, e) in zip(reqs, res):
        r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
        assert r == r0 == 0 and np.array_equal(e, e0)


def test_encode_tb_errors(eng):
    d = np.zeros(2000, np.uint8)
    assert eng.encode_tb(1000, 0, 0, 3000, d)[0] == -1            # Qm == 0
    assert eng.encode_tb(6152, 2, 0, 30000, d)[0] == -1           # filler bits
    assert eng.encode_tb(12216, 2, 0, 30000, d, max_cb=1)[0] == -1  # C > max_cb
    assert eng.encode_tb(1000, 2, 4, 3000, d)[0] == -1            # rv > 3
    assert eng.encode_tb(1000, 2, 0, 3000, None)[0] == -2         # no payload
    r, e = eng.encode_tb(0, 2, 0, 64, d)
    assert r == 0 and not e.any()                                 # tbs == 0: nothing written


def test_encode_then_decode_round_trip(sb, eng):
    """the engine's own encoder feeds its decoder: hard-decision LLRs of the e-bits decode back to the payload"""
    tbs, Qm = 75376, 6
    G = Qm * 20000
    data = np.random.default_rng(5).integers(0, 256, tbs // 8, dtype=np.uint8)
    r, e = eng.encode_tb(tbs, Qm, 0, G, data)
    assert r == 0
    llr = ((np.unpackbits(e)[:G].astype(np.int16) * 2 - 1) * 30).astype(np.int16)
    tb = sb.TransportBlock(tbs=tbs)
    assert eng.decode_tb(tb, Qm, 0, llr, 8) == 0
    assert np.array_equal(tb.data[:tbs // 8], data)


# ------------------------------------------------------------------ UL-SCH channel de-interleaver
def _ri_positions(Qm, rows, nsymb, nbits):
    """RI lands on columns 1,4,7,10 from the last row upwards (36.212 table 5.2.2.8-1), Qm bits per symbol"""
    cols = [c for c in (1, 4, 7, 10) if c < nsymb]
    pos = []
    for m in range(nbits // Qm):
        r, c = rows - 1 - m // len(cols), cols[m % len(cols)]
        pos += [r * Qm + c * rows * Qm + k for k in range(Qm)]
    return pos


@pytest.mark.parametrize("Qm", [2, 4, 6])
def test_ulsch_deinterleave_matches_oracle(eng, o, Qm):
    for nprb, nsymb in [(1, 12), (6, 12), (25, 10), (100, 12), (100, 11)]:
        H = nprb * 12 * nsymb
        rows = H // nsymb
        q = np.random.default_rng(H * 7 + Qm).integers(-30000, 30000, H * Qm).astype(np.int16)
        for ri in ([], _ri_positions(Qm, rows, nsymb, 8 * Qm), _ri_positions(Qm, rows, nsymb, 4 * Qm) + [0, 1], list(range(Qm))):
            ret, g = eng.ulsch_deinterleave(q, Qm, H, nsymb, ri)
            assert ret == 0
            g0 = o.ulsch_deinterleave(q, Qm, H, nsymb, ri)
            nd = H * Qm - len(set(ri))
            assert np.array_equal(g[:nd], g0[:nd]), (nprb, nsymb, len(ri))
    assert eng.ulsch_deinterleave(q, Qm, 145, 12, [])[0] == -2   # H' not a multiple of the symbol count


def test_ulsch_deinterleave_golden(eng):
    t = np.load(os.path.join(G, "tx.npz"))
    for n, (Qm, H, nsymb, nri) in enumerate(t["dei_cases"]):
        ri = list(t["dei%d_ri" % n])
        ret, g = eng.ulsch_deinterleave(t["dei%d_q" % n], int(Qm), int(H), int(nsymb), ri)
        nd = int(H) * int(Qm) - len(ri)
        assert ret == 0 and np.array_equal(g[:nd], t["dei%d_g" % n][:nd])


def test_ulsch_decode_from_interleaved_llrs(sb, eng, o):
    """srsran_ulsch_decode's data path on the device: q_bits -> de-interleave -> (skip RI and CQI) -> decode_tb, against the
    oracle's de-interleaver followed by the oracle's decode_tb; two HARQ transmissions, several TBs in one submission"""
    cases = [(12216, 4, 25, 12, 8, 0), (75376, 6, 100, 12, 24, 36), (2984, 2, 15, 11, 0, 0), (36696, 6, 50, 12, 12, 0)]
    tbl = [sb.TransportBlock(c[0]) for c in cases]
    st = [None] * len(cases)
    for rv in (0, 2):
        reqs, exp = [], []
        for n, (tbs, Qm, nprb, nsymb, nri_sym, ncqi_sym) in enumerate(cases):
            H = nprb * 12 * nsymb
            rows = H // nsymb
            ri = _ri_positions(Qm, rows, nsymb, nri_sym * Qm)
            e_off = ncqi_sym * Qm
            Gbits = H * Qm - len(ri) - e_off
            _, e = vecgen.make_tb(tbs, Gbits, Qm, rv, 3.0, 300 + n, scale=100)
            rng = np.random.default_rng(40 + n + rv)
            g = np.concatenate([rng.integers(-500, 500, e_off).astype(np.int16), e])
            # interleave: inverse of the reference's de-interleaver on the non-RI positions, noise on the RI positions
            q = rng.integers(-500, 500, H * Qm).astype(np.int16)
            # build the permutation explicitly (positions in (row, col, bit) scan order, RI skipped)
            is_ri = np.zeros(H * Qm, bool); is_ri[ri] = True
            j, i, k = np.meshgrid(np.arange(rows), np.arange(nsymb), np.arange(Qm), indexing="ij")
            p = (j * Qm + i * rows * Qm + k).reshape(-1)
            p = p[~is_ri[p]]
            q[p] = g[:len(p)]
            g_ref = o.ulsch_deinterleave(q, Qm, H, nsymb, ri)
            res = o.decode_tb(tbs, Qm, rv, g_ref[e_off:e_off + Gbits], 8, st[n])
            st[n] = res["state"]
            exp.append((res, g_ref, e_off))
            reqs.append((tbl[n], Qm, rv, q, H, nsymb, Gbits, ri, e_off, e_off))
        for tb in tbl:
            tb.data[:] = 0
        assert eng.ulsch_decode_batch(reqs, 8) == 0
        for n, (res, g_ref, e_off) in enumerate(exp):
            tb = tbl[n]
            Cn = res["seg"]["C"]
            assert tb.ret == res["ret"]
            assert np.array_equal(tb.cb_noi[:Cn], res["cb_noi"][:Cn])
            assert np.array_equal(tb.cb_crc[:Cn], res["state"]["cb_crc"][:Cn])
            assert np.array_equal(tb.buffer_f[:Cn], res["state"]["buffer_f"][:Cn])
            if res["ret"] == 0:
                assert np.array_equal(tb.data[:cases[n][0] // 8], res["data"][:cases[n][0] // 8])
            if e_off:
                assert np.array_equal(tb.g_bits[:e_off], g_ref[:e_off])
    assert all(tb.ret == 0 for tb in tbl)


# ------------------------------------------------------------------ device-resident, pipelined submissions (bench.py's path)
def test_device_resident_plans_pipelined(sb, eng, o):
    """srsb200_tdec_plan_uniform + srsb200_tdec_run_plan_dev: two plans used alternately (their submissions overlap on the
    GPU), the same plan re-submitted back to back (serialised), results after srsb200_engine_sync / _flush bit-exact"""
    import torch
    K, n = 1024, 2112   # 33 groups: the submission is split over sub-streams and joined lazily
    dev = torch.device("cuda", 0)
    sets = []
    for seed in (1, 2, 3):
        _, l48 = vecgen.make_cb_batch(K, 48, 1.0, 700 + seed)
        _, out, noi, ok = o.tdec_batch(K, l48, 8, True, nthreads=4)
        reps = -(-n // 48)
        llr = np.tile(l48, (reps, 1))[:n].copy()
        sets.append((torch.from_numpy(llr).to(dev), np.tile(out, (reps, 1))[:n], np.tile(noi, reps)[:n], np.tile(ok, reps)[:n]))
    plans = [eng.plan_uniform(n, K, sb.CRC_24B) for _ in range(2)]
    bufs = [(torch.zeros((n, K // 8), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev),
             torch.zeros(n, dtype=torch.uint8, device=dev)) for _ in range(6)]
    order = [(0, 0), (1, 1), (0, 2), (1, 0), (1, 1), (0, 2)]   # (plan, input set); submissions 3 and 4 reuse plan 1 back to back
    for b, (pi, si) in zip(bufs, order):
        eng.run_plan_dev(plans[pi], sets[si][0].data_ptr(), 8, 2, True, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr())
    eng.flush()
    eng.sync()
    for b, (pi, si) in zip(bufs, order):
        assert np.array_equal(b[1].cpu().numpy(), sets[si][2])
        assert np.array_equal(b[2].cpu().numpy(), sets[si][3])
        assert np.array_equal(b[0].cpu().numpy(), sets[si][1])
    for p in plans:
        eng.plan_destroy(p)


# ------------------------------------------------------------------ edge cases
def test_empty_submissions(sb, eng):
    """n == 0 everywhere: success, nothing launched, nothing written"""
    import ctypes as C
    L = sb.lib()
    n0 = eng.launch_count
    assert L.srsb200_tdec_batch(eng.handle, 0, None, None, None, None, 0, 8, 2, 1, None, None, 0, None, None) == 0
    assert L.srsb200_decode_tb_batch(eng.handle, None, 0, 8) == 0
    assert L.srsb200_encode_tb_batch(eng.handle, None, 0) == 0
    assert eng.launch_count == n0
    out, noi, ok = eng.tdec_batch(np.zeros(0, np.uint32), [], 8)
    assert len(noi) == 0


def test_largest_transport_block(sb, eng, o):
    """2-layer 100-PRB TBS 149776: C = 25 code blocks of K = 6016, decode over two HARQ transmissions and
    encode, against the oracle"""
    tbs, Qm = 149776, 6
    G = Qm * 28800
    _, seg = o.cbsegm(tbs)
    assert seg["C"] == 25 and seg["K1"] == 6016 and seg["F"] == 0
    tb = sb.TransportBlock(tbs)
    st = None
    for rv in (0, 2):
        _, e = vecgen.make_tb(tbs, G, Qm, rv, 3.5, 123, scale=700)
        res = o.decode_tb(tbs, Qm, rv, e, 8, st)
        st = res["state"]
        tb.data[:] = 0
        assert eng.decode_tb(tb, Qm, rv, e, 8) == res["ret"]
        assert np.array_equal(tb.cb_noi[:25], res["cb_noi"][:25]) and np.array_equal(tb.cb_crc[:25], st["cb_crc"][:25])
        assert np.array_equal(tb.buffer_f[:25], st["buffer_f"][:25])
        if res["ret"] == 0:
            assert np.array_equal(tb.data[:tbs // 8], res["data"][:tbs // 8])
    assert res["ret"] == 0
    data = np.random.default_rng(0).integers(0, 256, tbs // 8, dtype=np.uint8)
    r0, e0 = o.encode_tb(tbs, Qm, 1, G, data)
    r1, e1 = eng.encode_tb(tbs, Qm, 1, G, data)
    assert r0 == r1 == 0 and np.array_equal(e0, e1)


def test_every_block_size_one_submission(sb, eng, o):
    """all 188 LTE sizes, three blocks each (two clean, one hopeless), one mixed submission: bits, counts, verdicts"""
    Ks, llrs = [], []
    for idx in range(188):
        K = o.cbsize(idx)
        for j in range(3):
            _, l = vecgen.make_cb(K, 2.5 if j < 2 else -6.0, 9000 + 3 * idx + j)
            Ks.append(K)
            llrs.append(l)
    out, noi, ok = eng.tdec_batch(np.array(Ks, np.uint32), llrs, 6, early_stop=True)
    for i, K in enumerate(Ks):
        _, oo, on, ook = o.tdec_batch(K, llrs[i][None, :], 6, True)
        assert on[0] == noi[i] and ook[0] == ok[i] and np.array_equal(oo[0], out[i]), (K, i % 3)


def test_page_locked_caller_buffers(sb, eng, o):
    """e-bits and decoded bytes in page-locked memory from srsb200_host_alloc (contiguous, so the uploads merge into one DMA
    and the small read-backs ride the gather-copy kernel) give the same results as pageable numpy buffers"""
    import ctypes as C
    L = sb.lib()
    L.srsb200_host_alloc.restype = C.c_void_p
    L.srsb200_host_alloc.argtypes = [C.c_size_t]
    L.srsb200_host_free.argtypes = [C.c_void_p]
    cases = [(12216, 19200, 6), (6200, 9000, 4), (75376, 86400, 6), (2984, 4000, 2)]
    tot = sum(c[1] for c in cases)
    p_e = L.srsb200_host_alloc(tot * 2)
    p_d = L.srsb200_host_alloc(len(cases) * (13 * 768 + 8))
    assert p_e and p_d
    try:
        e_all = np.ctypeslib.as_array((C.c_int16 * tot).from_address(p_e))
        d_all = np.ctypeslib.as_array((C.c_uint8 * (len(cases) * (13 * 768 + 8))).from_address(p_d))
        reqs, exp, off = [], [], 0
        for n, (tbs, Gb, Qm) in enumerate(cases):
            _, e = vecgen.make_tb(tbs, Gb, Qm, 0, 3.0, 600 + n)
            e_all[off:off + Gb] = e
            tb = sb.TransportBlock(tbs)
            tb.data = d_all[n * (13 * 768 + 8):(n + 1) * (13 * 768 + 8)]
            tb.data[:] = 0
            reqs.append((tb, Qm, 0, e_all[off:off + Gb]))
            exp.append(o.decode_tb(tbs, Qm, 0, e, 8))
            off += Gb
        assert eng.decode_tb_batch(reqs, 8) == 0
        for (tb, _, _, _), r, (tbs, _, _) in zip(reqs, exp, cases):
            Cn = r["seg"]["C"]
            assert tb.ret == r["ret"] and np.array_equal(tb.cb_noi[:Cn], r["cb_noi"][:Cn])
            assert np.array_equal(tb.buffer_f[:Cn], r["state"]["buffer_f"][:Cn])
            if r["ret"] == 0:
                assert np.array_equal(tb.data[:tbs // 8], r["data"][:tbs // 8])
    finally:
        L.srsb200_host_free(p_e)
        L.srsb200_host_free(p_d)
    assert L.srsb200_host_register(None, 16) == -2


def test_randomized_submissions(sb, eng, o):
    """seeded sweep over submission shapes: random sets of block sizes and counts (partial groups, one block, > 64 equal
    blocks), CRC kind per block, early stop on / off, 1..10 half-iterations, good and hopeless channels"""
    rng = np.random.default_rng(20240 + int(os.environ.get("SRSB200_FUZZ_SEED", "0")))   # more seeds: tools/fuzz.sh
    sizes = [o.cbsize(i) for i in range(188)]
    small = [k for k in sizes if k <= 1024]
    for case in range(24):
        nk = int(rng.integers(1, 5))
        Ks, llrs, kinds = [], [], []
        for _ in range(nk):
            K = int(rng.choice(small if rng.random() < 0.8 else sizes))
            cnt = int(rng.choice([1, 2, 3, 31, 33, 64, 65, 70])) if K <= 256 else int(rng.integers(1, 6))
            eb = float(rng.choice([-3.0, 0.5, 1.5, 4.0]))
            kind = int(rng.integers(0, 3))
            for j in range(cnt):
                _, l = vecgen.make_cb(K, eb, int(rng.integers(1 << 30)), scale=int(rng.choice([30, 100, 400])), with_crc=(kind != 0))
                Ks.append(K); llrs.append(l); kinds.append(kind)
        max_iter = int(rng.integers(1, 11))
        early = bool(rng.integers(0, 2))
        out, noi, ok = eng.tdec_batch(np.array(Ks, np.uint32), llrs, max_iter, early_stop=early, crc_kind=np.array(kinds, np.uint8))
        for i in rng.permutation(len(Ks))[:12]:
            K = Ks[i]
            ref_out = o.tdec_run_all(K, llrs[i], max_iter) if not early or kinds[i] == 0 else None
            if ref_out is not None:
                # no early stop (or no CRC to stop on): all max_iter half-iterations run
                assert noi[i] == max_iter and np.array_equal(out[i], ref_out), (case, K, max_iter, early, kinds[i])
            else:
                hard = o.tdec_trace(K, llrs[i], max_iter)
                poly = ol.CRC24A if kinds[i] == 1 else ol.CRC24B
                stop = max_iter
                for it in range(1, max_iter + 1):
                    if it >= 2 and o.crc_bytes(poly, hard[it - 1], K) == 0:
                        stop = it
                        break
                assert noi[i] == stop and np.array_equal(out[i], hard[stop - 1]), (case, K, max_iter, kinds[i], noi[i], stop)


def test_randomized_transport_block_batches(sb, eng, o):
    """seeded sweep over transport-block submissions: random TBS / Qm / G / channel per block, 1-6 blocks per submission,
    HARQ state carried over three transmissions (rv 0, 2, 1) so that cached, failing and passing code blocks mix"""
    rng = np.random.default_rng(777 + int(os.environ.get("SRSB200_FUZZ_SEED", "0")))
    pool = [16, 40, 1000, 2984, 6120, 6200, 12216, 36696, 75376]
    for rnd in range(6):
        n = int(rng.integers(1, 7))
        specs = []
        for i in range(n):
            tbs = int(rng.choice(pool))
            Qm = int(rng.choice([2, 4, 6]))
            G = Qm * int(rng.integers(max(tbs // Qm // 2, 24), 2 * tbs // Qm + 60))
            specs.append((tbs, Qm, G, float(rng.choice([-1.0, 0.5, 2.0, 5.0])), int(rng.integers(1 << 30))))
        tbl = [sb.TransportBlock(s[0]) for s in specs]
        st = [None] * n
        for rv in (0, 2, 1):
            reqs, exp = [], []
            for i, (tbs, Qm, G, eb, seed) in enumerate(specs):
                _, e = vecgen.make_tb(tbs, G, Qm, rv, eb, seed, scale=100 if Qm < 6 else 400)
                res = o.decode_tb(tbs, Qm, rv, e, 7, st[i])
                st[i] = res["state"]
                exp.append(res)
                tbl[i].data[:] = 0
                reqs.append((tbl[i], Qm, rv, e))
            assert eng.decode_tb_batch(reqs, 7) == 0
            for tb, res in zip(tbl, exp):
                assert tb.ret == res["ret"] and int(tb.tb_crc[0]) == res["tb_crc"]
                C = res["seg"]["C"]
                assert np.array_equal(tb.cb_noi[:C], res["cb_noi"][:C]) and np.array_equal(tb.cb_crc[:C], res["state"]["cb_crc"][:C])
                assert np.array_equal(tb.buffer_f[:C], res["state"]["buffer_f"][:C]) and np.array_equal(tb.sb_data[:C], res["state"]["sb_data"][:C])
                if res["ret"] == 0:
                    assert np.array_equal(tb.data[:tb.tbs // 8], res["data"][:tb.tbs // 8])


@pytest.mark.parametrize("tbs,G,Qm", [(12216, 19200, 6), (75376, 86400, 6), (6120, 14400, 2), (40, 300, 2), (40, 72000, 2)])
def test_decode_tb_with_device_descrambling(sb, eng, o, tbs, G, Qm):
    """scrambled e-bits + c_init in, descrambling fused into the rate de-matcher (Gold sequence by GF(2) jump-ahead):
    identical to descrambling with the oracle's srsran_sequence_apply_s restatement first; (40, 72000) exercises repetition
    far beyond the staged part of the sequence"""
    c_init = (0x4321 << 14) + (1 << 13) + (6 << 9) + 77
    st = None
    tb = sb.TransportBlock(tbs)
    for rv in (0, 2):
        _, e = vecgen.make_tb(tbs, G, Qm, rv, 1.5, 910 + tbs, scale=100)
        e[0] = -32768                                     # its negation wraps
        scr = o.sequence_apply_s(e, c_init)              # what the demodulator hands over (scrambling is an involution)
        res = o.decode_tb(tbs, Qm, rv, o.sequence_apply_s(scr, c_init), 8, st)
        st = res["state"]
        tb.data[:] = 0
        assert eng.decode_tb(tb, Qm, rv, scr, 8, c_init=c_init) == res["ret"]
        _check_tb(res, tb, st)


def test_bench_path_vs_oracle(sb, o):
    """The path bench.py times: device-resident LLRs, srsb200_tdec_plan_uniform(8192+ blocks of K=6144) - i.e. 128+ full
    groups, the big-batch kernel instantiations - two plans used alternately so that consecutive submissions are pipelined
    on the engine's two lanes. Every block of 5 whole groups (320 blocks, one per range of the submission plus the last)
    is compared with the oracle: hard bytes, half-iteration counts, CRC verdicts."""
    import torch
    K, n = 6144, 8192 + 64 * 3 + 17  # a partial last group as well
    dev = torch.device("cuda", 0)
    e = sb.Engine(0)
    try:
        bits, llr16 = vecgen.make_cb_batch(K, 16, 1.5, 4242)
        coded = np.stack([o.encode(b) for b in bits])
        rng = np.random.default_rng(99)
        s = 2.0 * np.tile(coded, (n // 16 + 1, 1))[:n].astype(np.float64) - 1.0
        big = vecgen.quantise(s + vecgen.sigma_for(1.5) * rng.standard_normal(s.shape), 100)
        d_llr = torch.from_numpy(big).to(dev)
        plans = [e.plan_uniform(n, K, sb.CRC_24B) for _ in range(2)]
        outs = [(torch.zeros((n, K // 8), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev),
                 torch.zeros(n, dtype=torch.uint8, device=dev)) for _ in range(2)]
        for step in range(6):
            i = step % 2
            e.run_plan_dev(plans[i], d_llr.data_ptr(), 8, 2, True, outs[i][0].data_ptr(), outs[i][1].data_ptr(), outs[i][2].data_ptr())
        e.sync()
        torch.cuda.synchronize()
        for a, b in zip(outs[0], outs[1]):
            assert torch.equal(a, b)
        out, noi, ok = (t.cpu().numpy() for t in outs[0])
        ngroups = (n + 63) // 64
        sample = []
        for g in (0, ngroups // 4 + 1, ngroups // 2 + 2, 3 * ngroups // 4 + 3, ngroups - 2, ngroups - 1):
            sample += list(range(64 * g, min(n, 64 * g + 64)))
        sample = np.array(sample)
        assert len(sample) >= 320
        _, oo, on, ook = o.tdec_batch(K, big[sample], 8, True, nthreads=os.cpu_count() or 4)
        assert (on == noi[sample]).all() and (ook == ok[sample]).all() and (oo == out[sample]).all()
        assert ok.mean() > 0.97 and noi.min() >= 2
        for p in plans:
            e.plan_destroy(p)
    finally:
        e.close()


@pytest.mark.parametrize("mix", ["few_hard", "all_hard", "all_easy"])
def test_regrouping_of_unfinished_blocks_vs_oracle(sb, o, mix):
    """Plans of one block size with 64+ groups pack the unfinished code blocks of a range into fresh groups once few enough are
    left (regroup_plan_kernel). Mixed difficulty: most blocks finish after two or three half-iterations, one in ten runs long or
    never passes - every block of the batch is compared with the oracle (bytes, half-iteration count, CRC verdict). `all_hard`
    never regroups (too many survivors), `all_easy` has nothing left at the first regrouping point."""
    import torch
    K, n = 1024, 64 * 70 + 9
    dev = torch.device("cuda", 0)
    e = sb.Engine(0)
    try:
        rng = np.random.default_rng(31)
        bits, _ = vecgen.make_cb_batch(K, 24, 3.0, 777)
        coded = np.stack([o.encode(b) for b in bits])
        s = 2.0 * coded[rng.integers(0, 24, n)].astype(np.float64) - 1.0
        hard = {"few_hard": rng.random(n) < 0.1, "all_hard": np.ones(n, bool), "all_easy": np.zeros(n, bool)}[mix]
        eb = np.where(hard, rng.choice([0.2, 0.9, 1.3], n), 4.5)
        sig = np.array([vecgen.sigma_for(x) for x in eb])[:, None]
        llr = vecgen.quantise(s + sig * rng.standard_normal(s.shape), 60)
        d_llr = torch.from_numpy(llr).to(dev)
        plans = [e.plan_uniform(n, K, sb.CRC_24B) for _ in range(2)]
        outs = [(torch.zeros((n, K // 8), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev),
                 torch.zeros(n, dtype=torch.uint8, device=dev)) for _ in range(2)]
        for step in range(4):  # the second use of a plan starts from the slots the first one left behind
            i = step % 2
            e.run_plan_dev(plans[i], d_llr.data_ptr(), 8, 2, True, outs[i][0].data_ptr(), outs[i][1].data_ptr(), outs[i][2].data_ptr())
        e.sync()
        torch.cuda.synchronize()
        _, oo, on, ook = o.tdec_batch(K, llr, 8, True, nthreads=os.cpu_count() or 4)
        for out_t, noi_t, ok_t in outs:
            out, noi, ok = out_t.cpu().numpy(), noi_t.cpu().numpy(), ok_t.cpu().numpy()
            assert (on == noi).all(), np.flatnonzero(on != noi)[:10]
            assert (ook == ok).all()
            assert (oo == out).all(), np.flatnonzero((oo != out).any(axis=1))[:10]
        pts = [e.plan_regroup_points(p) for p in plans]
        if mix == "few_hard":
            assert (on[~hard] <= 4).mean() > 0.9 and (on[hard] >= 6).mean() > 0.3  # the case the test is about
            assert all(sum(1 for x in pp if x == 4) >= 3 for pp in pts), pts       # (nearly) every range regrouped at the first point
        else:
            assert all(not any(pp) for pp in pts), pts
        for p in plans:
            e.plan_destroy(p)
    finally:
        e.close()


def test_randomized_regrouping(sb, o):
    """seeded sweep over plans that may regroup: block size, number of groups (64..130, partial last group), share and difficulty
    of the hard blocks, half-iteration limit - every block against the oracle, two plans used alternately"""
    import torch
    rng = np.random.default_rng(4711 + int(os.environ.get("SRSB200_FUZZ_SEED", "0")))   # more seeds: tools/fuzz.sh
    dev = torch.device("cuda", 0)
    e = sb.Engine(0)
    try:
        for case in range(4):
            K = int(rng.choice([40, 512, 1024, 2048]))
            n = 64 * int(rng.integers(64, 131)) - int(rng.integers(0, 64))
            max_iter = int(rng.integers(6, 11))
            bits, _ = vecgen.make_cb_batch(K, 16, 3.0, int(rng.integers(1 << 20)))
            coded = np.stack([o.encode(b) for b in bits])
            s_ = 2.0 * coded[rng.integers(0, 16, n)].astype(np.float64) - 1.0
            share = float(rng.choice([0.02, 0.1, 0.3, 0.6]))
            hard = rng.random(n) < share
            eb = np.where(hard, rng.choice([-1.0, 0.4, 1.0, 1.6], n), rng.choice([3.5, 5.0], n))
            sig = np.array([vecgen.sigma_for(x) for x in eb])[:, None]
            llr = vecgen.quantise(s_ + sig * rng.standard_normal(s_.shape), int(rng.choice([40, 100, 300])))
            d_llr = torch.from_numpy(llr).to(dev)
            plans = [e.plan_uniform(n, K, sb.CRC_24B) for _ in range(2)]
            outs = [(torch.zeros((n, K // 8), dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev),
                     torch.zeros(n, dtype=torch.uint8, device=dev)) for _ in range(2)]
            for step in range(4):
                i = step % 2
                e.run_plan_dev(plans[i], d_llr.data_ptr(), max_iter, 2, True, outs[i][0].data_ptr(), outs[i][1].data_ptr(), outs[i][2].data_ptr())
            e.sync()
            torch.cuda.synchronize()
            _, oo, on, ook = o.tdec_batch(K, llr, max_iter, True, nthreads=os.cpu_count() or 4)
            for out_t, noi_t, ok_t in outs:
                out, noi, ok = out_t.cpu().numpy(), noi_t.cpu().numpy(), ok_t.cpu().numpy()
                assert (on == noi).all() and (ook == ok).all() and (oo == out).all(), (case, K, n, max_iter, share, [e.plan_regroup_points(p) for p in plans])
            for p in plans:
                e.plan_destroy(p)
    finally:
        e.close()


# ---------------------------------------------------------------- robustness of the C ABI
def test_decode_tb_batch_error_paths_fault_injection(sb, o):
    """srsb200_engine_inject_alloc_failure: whichever scratch request of a transport-block submission fails, the call returns an
    error with every queued block's ret = -1, nothing is left in flight, and the very next submission is bit-exact again"""
    e = sb.Engine(0)
    try:
        cases = [(36696, 43200, 6, 1.0), (6200, 9000, 4, 0.5), (12216, 19200, 6, 2.5)]
        es = [vecgen.make_tb(tbs, Gb, Qm, 0, eb, 400 + i)[1] for i, (tbs, Gb, Qm, eb) in enumerate(cases)]
        exp = [o.decode_tb(tbs, Qm, 0, ev, 8) for (tbs, Gb, Qm, eb), ev in zip(cases, es)]
        failed = 0
        for nth in range(1, 12):
            tbo = [sb.TransportBlock(c[0]) for c in cases]
            reqs = [(tb, c[2], 0, ev) for tb, c, ev in zip(tbo, cases, es)]
            e.inject_alloc_failure(nth)
            ret = e.decode_tb_batch(reqs, 8)
            e.inject_alloc_failure(0)
            if ret != 0:
                failed += 1
                assert all(tb.ret == -1 for tb in tbo), (nth, [tb.ret for tb in tbo])
            tbo = [sb.TransportBlock(c[0]) for c in cases]
            reqs = [(tb, c[2], 0, ev) for tb, c, ev in zip(tbo, cases, es)]
            assert e.decode_tb_batch(reqs, 8) == 0
            for tb, r in zip(tbo, exp):
                _check_tb(r, tb, r["state"])
        assert failed >= 5
    finally:
        e.close()


def test_decode_tb_batch_per_tb_iteration_limits(sb, eng, o):
    """transport blocks of one submission with different half-iteration limits (one srsran_sch_t each): every block
    must stop at ITS limit, as decode_tb called once per TB does"""
    cases = [(36696, 43200, 6, 0.3, 2), (36696, 43200, 6, 0.3, 8), (6200, 9000, 4, 0.2, 3), (12216, 19200, 6, 0.6, 5)]
    tbo, reqs, exp = [], [], []
    for i, (tbs, Gb, Qm, eb, lim) in enumerate(cases):
        _, ev = vecgen.make_tb(tbs, Gb, Qm, 0, eb, 500 + i // 2)
        tb = sb.TransportBlock(tbs)
        tbo.append(tb); reqs.append((tb, Qm, 0, ev)); exp.append(o.decode_tb(tbs, Qm, 0, ev, lim))
    assert eng.decode_tb_batch(reqs, 8, limits=[c[4] for c in cases]) == 0
    for tb, r in zip(tbo, exp):
        _check_tb(r, tb, r["state"])
    assert exp[0]["cb_noi"].max() == 2 and exp[1]["cb_noi"].max() > 2   # same e-bits, different limits: they do differ


def test_tdec_batch_scattered_output_keeps_gaps(sb, eng, o):
    """srsb200_tdec_batch with non-contiguous output offsets writes exactly [out_offset[i], +K/8): the bytes between
    and after the blocks keep the caller's content"""
    import ctypes as C
    K = 512
    _, llr = vecgen.make_cb_batch(K, 3, 2.0, 808)
    _, oo, on, ook = o.tdec_batch(K, llr, 6, True)
    Ks = np.full(3, K, np.uint32); kinds = np.full(3, sb.CRC_24B, np.uint8)
    loff = (np.arange(3, dtype=np.uint64) * np.uint64(3 * K + 12))
    ooff = np.array([5, 100, 301], np.uint64)   # odd offsets, gaps
    out = np.full(400, 0xA5, np.uint8); noi = np.zeros(3, np.uint8); ok = np.zeros(3, np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    flat = np.ascontiguousarray(llr.reshape(-1))
    assert sb.lib().srsb200_tdec_batch(eng.handle, 3, vp(Ks), vp(kinds), vp(flat), vp(loff), len(flat), 6, 2, 1, vp(out), vp(ooff), len(out), vp(noi), vp(ok)) == 0
    mask = np.ones(400, bool)
    for i in range(3):
        assert (out[int(ooff[i]):int(ooff[i]) + K // 8] == oo[i]).all() and noi[i] == on[i] and ok[i] == ook[i]
        mask[int(ooff[i]):int(ooff[i]) + K // 8] = False
    assert (out[mask] == 0xA5).all()


def test_multi_device_dispatcher(sb, o):
    """srsb200_multi_*: transport blocks placed by owner key modulo the device count, flat batches split by sum(K); every result
    equals the oracle's whatever the number of devices; with two or more devices every one of them must have launched kernels"""
    import torch
    m = sb.Multi()
    try:
        nd = m.nof_devices
        assert nd == torch.cuda.device_count() >= 1
        cases = [(36696, 43200, 6, 1.0), (6200, 9000, 4, 0.5), (12216, 19200, 6, 2.5), (75376, 86400, 6, 6.0), (6120, 14400, 2, 3.0), (40, 300, 2, 3.0)]
        for round_ in range(2):   # second round: HARQ retransmission, every owner must land on the same device again
            rv = (0, 2)[round_]
            if round_ == 0:
                tbo = [sb.TransportBlock(c[0]) for c in cases]
                st = [None] * len(cases)
            reqs, exp = [], []
            for i, (tbs, Gb, Qm, eb) in enumerate(cases):
                _, ev = vecgen.make_tb(tbs, Gb, Qm, rv, eb - 1.5, 600 + i)
                reqs.append((tbo[i], Qm, rv, ev))
                r = o.decode_tb(tbs, Qm, rv, ev, 8, st[i]); st[i] = r["state"]; exp.append(r)
                tbo[i].data[:] = 0
            assert m.decode_tb_batch(reqs, 8, owners=list(range(len(cases)))) == 0, sb.lib().srsb200_last_error().decode()
            for tb, r, s_ in zip(tbo, exp, st):
                _check_tb(r, tb, s_)
        assert [m.device_of(i) for i in range(6)] == [i % nd for i in range(6)]
        K = 1024
        _, llr = vecgen.make_cb_batch(K, 70, 1.2, 4321)
        _, oo, on, ook = o.tdec_batch(K, llr, 7, True, nthreads=4)
        out, noi, ok = m.tdec_batch(K, llr, 7)
        assert (out == oo).all() and (noi == on).all() and (ok == ook).all()
        assert all(c > 0 for c in m.launch_counts())
    finally:
        m.close()


# ---------------------------------------------------------------- soft demodulation + descrambling on the device
QAM_NORM = {1: np.sqrt(2.0), 2: np.sqrt(10.0), 3: np.sqrt(42.0), 4: np.sqrt(170.0)}


def _modulate(bits, mod):
    """36.211 7.1 constellations from hard bits (test-signal generation only) -> complex64 symbols"""
    bps = (1, 2, 4, 6, 8)[mod]
    b = 1.0 - 2.0 * np.asarray(bits, np.float64).reshape(-1, bps)   # bit 0 -> +1
    if mod == 0:
        v = b[:, 0] / np.sqrt(2.0)
        return (v + 1j * v).astype(np.complex64)

    def axis(first):
        cols = list(range(first, bps, 2))
        level = [1, 2, 4, 8][:len(cols)][::-1]          # e.g. 64QAM: 4, 2, 1
        acc = np.full(len(b), float(level[-1]))
        for j in range(len(cols) - 1, 0, -1):           # innermost term first: (2 - (1 - 2 b4)) ...
            acc = level[j - 1] - b[:, cols[j]] * acc
        return b[:, cols[0]] * acc
    return ((axis(0) + 1j * axis(1)) / QAM_NORM[mod]).astype(np.complex64)


@pytest.mark.parametrize("mod", [0, 1, 2, 3, 4])
def test_demod_soft_demodulate_s(sb, eng, o, mod):
    rng = np.random.default_rng(50 + mod)
    for n in (1, 3, 4, 5, 8, 9, 17, 1201, 14400):
        for amp in (0.3, 1.0, 30.0, 200.0):
            s = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
            ret, llr = eng.demod_soft_demodulate_s(mod, s)
            assert ret == 0 and np.array_equal(llr, o.demod_soft_demodulate_s(mod, s)), (mod, n, amp)
    assert eng.demod_soft_demodulate_s(7, np.zeros(4, np.complex64))[0] == -1


@pytest.mark.parametrize("tbs,nre,mod,eb", [(75376, 14400, 3, 12.5), (36696, 7201, 3, 12.0), (12216, 4803, 2, 6.5), (6120, 7200, 1, 2.5), (75376, 11000, 4, 18.5)])
def test_decode_tb_from_symbols_downlink(sb, eng, o, tbs, nre, mod, eb):
    """pdsch.c:693-740 on the device: equalised symbols -> soft demodulation -> descrambling -> rate de-matching -> decode, against
    the oracle chain demod -> sequence_apply_s -> decode_tb on the same symbols (odd symbol counts exercise the scalar tails)"""
    Qm = (1, 2, 4, 6, 8)[mod]
    G = nre * Qm
    rng = np.random.default_rng(tbs + mod)
    c_init = (0x1234 << 14) | (3 << 9) | 77
    payload, e_clean = vecgen.make_tb(tbs, G, Qm, 0, 60.0, 5 + mod, scale=8)   # noiseless LLRs: sign = transmitted bit
    tx_bits = (e_clean > 0).astype(np.uint8)
    scr = o.sequence_apply_s(np.ones(G, np.int16), c_init) < 0
    sym = _modulate(tx_bits ^ scr.astype(np.uint8), mod)
    sigma = vecgen.sigma_for(eb, tbs / float(G)) / np.sqrt(float(Qm))   # per component, unit symbol energy: N0 / 2 = 1 / (2 Qm R Eb/N0)
    sym = (sym + sigma * (rng.standard_normal(nre) + 1j * rng.standard_normal(nre))).astype(np.complex64)
    llr = o.sequence_apply_s(o.demod_soft_demodulate_s(mod, sym), c_init)
    ref_res = o.decode_tb(tbs, Qm, 0, llr, 8)
    tb = sb.TransportBlock(tbs)
    assert eng.decode_tb_symbols(tb, Qm, 0, sym, mod, G, 8, c_init=c_init) == ref_res["ret"]
    _check_tb(ref_res, tb, ref_res["state"])
    assert ref_res["ret"] == 0 and np.array_equal(ref_res["data"][:tbs // 8], payload[:tbs // 8])


@pytest.mark.parametrize("tbs,nprb,nsymb,mod,ri", [(36696, 50, 12, 3, True), (12216, 25, 12, 2, False), (6120, 30, 11, 1, True)])
def test_decode_tb_from_symbols_uplink(sb, eng, o, tbs, nprb, nsymb, mod, ri):
    """pusch.c:418-455 on the device: symbols -> demodulation -> descrambling of the interleaved stream -> channel de-interleaver ->
    decode; the descrambled LLRs at the RI positions come back for the host-side UCI decoding"""
    Qm = (1, 2, 4, 6, 8)[mod]
    H = nprb * 12 * nsymb
    rows = H // nsymb
    c_init = (0x4321 << 14) | (7 << 9) | 301
    pos = []
    if ri:
        for n_ in range(8):
            r_ = rows - 1 - n_ // 4
            c_ = (1, 4, 7, 10)[n_ % 4]
            pos += [r_ * Qm + c_ * rows * Qm + k for k in range(Qm)]
    G = H * Qm - len(pos)
    rng = np.random.default_rng(tbs + 9)
    payload, e_clean = vecgen.make_tb(tbs, G, Qm, 0, 60.0, 8 + mod, scale=8)
    # interleave the clean bits the way the de-interleaver will undo it: build q so that deinterleave(q) = e
    idx = o.ulsch_deinterleave(np.arange(H * Qm, dtype=np.int32).astype(np.int16) * 0, Qm, H, nsymb, pos)   # shape probe
    q_bits = np.zeros(H * Qm, np.uint8)
    # (find the permutation with unique markers in two passes of 15-bit values)
    marks = np.arange(H * Qm, dtype=np.int64)
    lo = o.ulsch_deinterleave((marks & 0x7fff).astype(np.int16), Qm, H, nsymb, pos).astype(np.int64) & 0x7fff
    hi = o.ulsch_deinterleave((marks >> 15).astype(np.int16), Qm, H, nsymb, pos).astype(np.int64)
    src = (hi << 15) | lo           # g[r] = q[src[r]]
    q_bits[src[:G]] = (e_clean > 0).astype(np.uint8)
    scr = o.sequence_apply_s(np.ones(H * Qm, np.int16), c_init) < 0
    sym = _modulate(q_bits ^ scr.astype(np.uint8), mod)
    sigma = vecgen.sigma_for({1: 6.0, 2: 10.5, 3: 12.5}[mod], tbs / float(G)) / np.sqrt(float(Qm))
    sym = (sym + sigma * (rng.standard_normal(H) + 1j * rng.standard_normal(H))).astype(np.complex64)
    q = o.sequence_apply_s(o.demod_soft_demodulate_s(mod, sym), c_init)
    g = o.ulsch_deinterleave(q, Qm, H, nsymb, pos)
    ref_res = o.decode_tb(tbs, Qm, 0, g[:G], 8)
    tb = sb.TransportBlock(tbs)
    ret = eng.decode_tb_symbols(tb, Qm, 0, sym, mod, G, 8, c_init=c_init, ul=dict(H_prime_total=H, N_pusch_symbs=nsymb, ri_positions=pos),
                                q_gather=pos[:11] + [0, H * Qm - 1])
    assert ret == ref_res["ret"]
    _check_tb(ref_res, tb, ref_res["state"])
    exp = q[np.array(pos[:11] + [0, H * Qm - 1], np.int64)]
    assert np.array_equal(tb.q_gather_out[:len(exp)], exp)
    assert ref_res["ret"] == 0 and np.array_equal(ref_res["data"][:tbs // 8], payload[:tbs // 8])
    del idx
