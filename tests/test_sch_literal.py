"""The oracle against the LITERAL lib/src/phy/phch/sch.c of the reference (compiled into oracle/_ref together with the
UCI/CQI/convolutional code it links against): srsran_dlsch_encode2, srsran_dlsch_decode2, ulsch_deinterleave."""
import numpy as np
import pytest

import oracle_lib as ol
import vecgen

ref = ol.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


@pytest.mark.parametrize("tbs", [40, 1000, 6120, 6200, 12216, 36696, 75376])
@pytest.mark.parametrize("Qm,rv", [(2, 0), (4, 1), (6, 2), (2, 3), (6, 0)])
def test_encode_tb_vs_srsran_dlsch_encode2(tbs, Qm, rv):
    o = ol.oracle()
    _, seg = o.cbsegm(tbs)
    if seg["F"]:
        pytest.skip("filler bits")
    rng = np.random.default_rng(tbs + Qm + rv)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    for G in (Qm * ((tbs * 2) // Qm), Qm * ((tbs * 4 + 12 * seg["C"]) // Qm), Qm * (tbs // Qm // 2 + 7)):
        r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
        r1, e1 = ref.dlsch_encode(tbs, Qm, rv, G, data)
        assert r0 == r1 == 0
        assert np.array_equal(np.unpackbits(e0)[:G], np.unpackbits(e1)[:G])


def test_dlsch_encode2_error_codes():
    o = ol.oracle()
    d = np.zeros(2000, np.uint8)
    assert ref.dlsch_encode(6152, 2, 0, 30000, d)[0] == o.encode_tb(6152, 2, 0, 30000, d)[0] == -1   # filler bits


def _harq_case(tbs, Qm, G, kill, seed):
    """three transmissions (rv 0, 2, 0), high SNR; in the first one the LLRs of the code blocks in `kill` are replaced by noise so
    they certainly fail while the others certainly pass - no dependence on which decoder flavour runs"""
    o = ol.oracle()
    _, seg = o.cbsegm(tbs)
    Cn = seg["C"]
    payload, e0 = vecgen.make_tb(tbs, G, Qm, 0, 12.0, seed, scale=40)
    _, e2 = vecgen.make_tb(tbs, G, Qm, 2, 12.0, seed, scale=40, payload=np.unpackbits(payload)[:tbs])
    Gp = G // Qm
    gamma = Gp % Cn
    n_e = Qm * (Gp // Cn)
    e0 = e0.copy()
    for r in kill:
        E, rp = n_e, r * n_e
        if r > Cn - gamma:
            E = n_e + Qm
            rp = (Cn - gamma) * n_e + (r - (Cn - gamma)) * E
        e0[rp:rp + E] = np.random.default_rng(seed + r).integers(-40, 41, E)   # garbage (all-zero LLRs would decode to the all-zero codeword, CRC 0)
    _, e0b = vecgen.make_tb(tbs, G, Qm, 0, 12.0, seed + 1, scale=40, payload=np.unpackbits(payload)[:tbs])
    return payload, [(0, e0), (2, e2), (0, e0b)], seg


@pytest.mark.parametrize("tbs,Qm,G,kill", [(12216, 4, 4 * 4500, [1]), (36696, 6, 6 * 8000, [0, 3]), (75376, 6, 86400, [5, 12]), (6120, 2, 2 * 5000, [0]),
                                           (36696, 2, 2 * 30000, [])])
def test_decode_tb_loop_vs_srsran_dlsch_decode2(tbs, Qm, G, kill):
    """return code, data bytes, cb_crc / tb_crc flags of the restated decode_tb loop equal the literal srsran_dlsch_decode2
    (production AUTO decoder) across a HARQ retransmission with cached code blocks"""
    o = ol.oracle()
    payload, txs, seg = _harq_case(tbs, Qm, G, kill, 77 + tbs)
    h = ref.dlsch_rx_new()
    st = None
    decoded = False
    try:
        for rv, e in txs:
            a = o.decode_tb(tbs, Qm, rv, e, 8, st)
            st = a["state"]
            b = ref.dlsch_decode(h, tbs, Qm, rv, e, 8)
            assert a["ret"] == b["ret"]
            assert a["tb_crc"] == b["tb_crc"]
            assert np.array_equal(st["cb_crc"][:seg["C"]], b["cb_crc"][:seg["C"]])
            assert np.array_equal(a["data"][:tbs // 8], b["data"][:tbs // 8]) or a["ret"] != 0
            if a["ret"] == 0 and not decoded:
                decoded = True
                assert np.array_equal(a["data"][:tbs // 8], payload[:tbs // 8])
            # (a transmission AFTER the TB has passed skips every code block and returns the never-filled cache:
            #  sch.c:391-392,468-473 - the loop above still checks that both sides agree on that)
        assert decoded
    finally:
        ref.dlsch_rx_free(h)


@pytest.mark.parametrize("Qm", [2, 4, 6])
@pytest.mark.parametrize("nprb,nsymb", [(1, 12), (6, 12), (25, 10), (100, 12), (100, 11)])
def test_ulsch_deinterleave_vs_reference(Qm, nprb, nsymb):
    o = ol.oracle()
    H = nprb * 12 * nsymb
    rows = H // nsymb
    rng = np.random.default_rng(H + Qm)
    q = rng.integers(-30000, 30000, H * Qm).astype(np.int16)
    # no RI, then RI on the four RI columns (36.212 table 5.2.2.8-1 columns 1,4,7,10) of the last rows, then on position 0
    cases = [[]]
    ri_cols = [c for c in (1, 4, 7, 10) if c < nsymb]
    pos = []
    for n_ in range(min(8, rows)):
        r = rows - 1 - n_ // 4
        c = ri_cols[n_ % len(ri_cols)]
        pos += [r * Qm + c * rows * Qm + k for k in range(Qm)]
    cases.append(pos)
    cases.append(pos + [0, 1])
    for ri in cases:
        g0 = o.ulsch_deinterleave(q, Qm, H, nsymb, ri)
        g1 = ref.ulsch_deinterleave(q, Qm, H, nsymb, ri)
        n_data = H * Qm - len(set(ri))
        assert np.array_equal(g0[:n_data], g1[:n_data])


@pytest.mark.parametrize("tbs,Qm,nprb,nsymb,ri_len", [(2984, 2, 15, 12, 0), (12216, 4, 25, 12, 1), (36696, 6, 50, 12, 1), (6120, 2, 30, 11, 0)])
def test_ulsch_chain_literal(tbs, Qm, nprb, nsymb, ri_len):
    """srsran_ulsch_encode -> LLRs -> srsran_ulsch_decode (both literal) returns the payload; without RI the oracle chain
    (de-interleaver + decode_tb) agrees with it on the bytes"""
    o = ol.oracle()
    rng = np.random.default_rng(tbs)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    ret, qb = ref.ulsch_encode(tbs, Qm, 0, nsymb, nprb, data, ri_len, 1)
    assert ret >= 0   # srsran_ulsch_encode returns the number of RI/ACK bits it placed
    llr = ((qb.astype(np.int16) * 2 - 1) * 60).astype(np.int16)
    h = ref.dlsch_rx_new()
    try:
        r = ref.ulsch_decode(h, tbs, Qm, 0, nsymb, nprb, llr, 8, ri_len)
    finally:
        ref.dlsch_rx_free(h)
    assert r["ret"] == 0 and np.array_equal(r["data"][:tbs // 8], data)
    if not ri_len:   # (the RI value itself is UCI control decoding with the scrambling sequence: not this path)
        H = nprb * 12 * nsymb
        g = o.ulsch_deinterleave(llr, Qm, H, nsymb, [])
        a = o.decode_tb(tbs, Qm, 0, g, 8)
        assert a["ret"] == 0 and np.array_equal(a["data"][:tbs // 8], data)


# ---------------------------------------------------------------- 2 layers, second codeword, 256QAM (sch.c:580-609)
# (tbs, Qm of the modulation, G, kill, tb_idx, nof_layers, nof_tb): Nl = 2 whenever nof_layers != nof_tb
CW_CASES = [
    (149776, 6, 12 * 14400, [3, 24], 0, 2, 1),   # BASELINE config 3: 100 PRB 64QAM on two layers, one TB -> Qm * Nl = 12, C = 25, K = 6016
    (75376, 6, 86400, [12], 1, 2, 2),            # second codeword of a two-codeword grant (Nl = 1)
    (75376, 8, 8 * 11000, [0, 7], 0, 1, 1),      # 256QAM
    (97896, 8, 16 * 7200, [15], 1, 4, 2),        # 256QAM, four layers on two codewords -> Qm * Nl = 16, second codeword
    (36696, 4, 8 * 5500, [], 0, 2, 1),           # 16QAM on two layers -> 8
]


@pytest.mark.parametrize("tbs,Qm,G,kill,tb_idx,nof_layers,nof_tb", CW_CASES)
def test_decode_tb_layers_codewords_vs_srsran_dlsch_decode2(tbs, Qm, G, kill, tb_idx, nof_layers, nof_tb):
    o = ol.oracle()
    Qe = Qm * (2 if nof_layers != nof_tb else 1)
    payload, txs, seg = _harq_case(tbs, Qe, G, kill, 91 + tbs + tb_idx)
    if seg["C"] > 16:
        # the stock 110-PRB soft buffer holds 16 code blocks: decode_tb refuses (sch.c:541-545; the product mirrors it through
        # srsb200_tb_t.max_cb, tests/test_gpu_parity.py::test_decode_tb_layers_codewords)
        h = ref.dlsch_rx_new()
        try:
            assert ref.dlsch_decode_cw(h, tbs, Qm, 0, txs[0][1], 8, tb_idx, nof_layers, nof_tb)["ret"] == -2
        finally:
            ref.dlsch_rx_free(h)
    h = ref.dlsch_rx_new_guru(max(seg["C"], 16))
    st, decoded = None, False
    try:
        for rv, e in txs:
            a = o.decode_tb(tbs, Qe, rv, e, 8, st)
            st = a["state"]
            b = ref.dlsch_decode_cw(h, tbs, Qm, rv, e, 8, tb_idx, nof_layers, nof_tb)
            assert a["ret"] == b["ret"] and a["tb_crc"] == b["tb_crc"]
            assert np.array_equal(st["cb_crc"][:seg["C"]], b["cb_crc"][:seg["C"]])
            assert np.array_equal(a["data"][:tbs // 8], b["data"][:tbs // 8]) or a["ret"] != 0
            if a["ret"] == 0 and not decoded:
                decoded = True
                assert np.array_equal(a["data"][:tbs // 8], payload[:tbs // 8])
        assert decoded
    finally:
        ref.dlsch_rx_free(h)


@pytest.mark.parametrize("tbs,Qm,G,kill,tb_idx,nof_layers,nof_tb", CW_CASES)
@pytest.mark.parametrize("rv", [0, 2])
def test_encode_tb_layers_codewords_vs_srsran_dlsch_encode2(tbs, Qm, G, kill, tb_idx, nof_layers, nof_tb, rv):
    o = ol.oracle()
    Qe = Qm * (2 if nof_layers != nof_tb else 1)
    data = np.random.default_rng(tbs + rv).integers(0, 256, tbs // 8, dtype=np.uint8)
    r0, e0 = o.encode_tb(tbs, Qe, rv, G, data)
    r1, e1 = ref.dlsch_encode_cw(tbs, Qm, rv, G, data, tb_idx, nof_layers, nof_tb)
    assert r0 == r1 == 0
    assert np.array_equal(np.unpackbits(e0)[:G], np.unpackbits(e1)[:G])


@pytest.mark.parametrize("tbs,Qm,rv,G", [(12216, 6, 2, 19200), (75376, 6, 3, 86400), (6120, 2, 1, 9000), (40, 2, 2, 120)])
def test_encode_retransmission_without_payload_literal(tbs, Qm, rv, G):
    """sch.c:305: data == NULL re-reads the circular buffers of the previous call = the same bits as encoding the payload at that rv"""
    o = ol.oracle()
    data = np.random.default_rng(tbs + 5).integers(0, 256, tbs // 8, dtype=np.uint8)
    r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
    r1, e1 = ref.dlsch_encode_retx_null(tbs, Qm, rv, G, data)
    nb = Qm * (G // Qm)
    assert r0 == r1 == 0 and np.array_equal(np.unpackbits(e0)[:nb], np.unpackbits(e1)[:nb])
