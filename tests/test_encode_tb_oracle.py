"""Transport-block encode: the clean-room oracle pinned against the reference's own primitives
(srsran_tcod_encode_lut + srsran_rm_turbo_tx_lut composed as encode_tb_off, sch.c:240-358) compiled in oracle/_ref."""
import numpy as np
import pytest

import oracle_lib as ol

ref = ol.ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")

# tbs: 1 CB (40..6120), several equal CBs (standard TBS), C2 > 0 with F == 0 (non-standard, exercises the K2-first order)
TBS_CASES = [16, 40, 104, 1000, 2984, 6120, 6200, 12216, 36696, 75376, 6144 * 2 - 48 - 24 - 64]


def _written_bits(Qm, G):
    return Qm * (G // Qm)


@needs_ref
@pytest.mark.parametrize("tbs", TBS_CASES)
@pytest.mark.parametrize("Qm,rv", [(2, 0), (4, 1), (6, 2), (2, 3), (6, 0)])
def test_encode_tb_oracle_vs_ref(tbs, Qm, rv):
    o = ol.oracle()
    ret, seg = o.cbsegm(tbs)
    if ret or seg["F"]:
        pytest.skip("needs filler bits")
    rng = np.random.default_rng(tbs * 7 + Qm + rv)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    for G in (Qm * ((tbs * 2) // Qm), Qm * ((tbs * 4 + 12 * seg["C"]) // Qm) + 1, Qm * (tbs // Qm // 2 + 7)):
        r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
        r1, e1 = ref.encode_tb(tbs, Qm, rv, G, data)
        assert r0 == r1 == 0
        nb = _written_bits(Qm, G)
        b0 = np.unpackbits(e0)[:nb]
        b1 = np.unpackbits(e1)[:nb]
        assert np.array_equal(b0, b1), f"tbs={tbs} G={G}: first diff at {np.flatnonzero(b0 != b1)[:5]}"


@needs_ref
def test_encode_tb_c2_case_exists():
    """at least one case above has C2 > 0 and F == 0, so the 'K2 blocks first' order is really exercised"""
    o = ol.oracle()
    found = False
    for tbs in range(6200, 30000, 8):
        ret, seg = o.cbsegm(tbs)
        if ret == 0 and seg["F"] == 0 and seg["C2"] > 0:
            found = True
            rng = np.random.default_rng(tbs)
            data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            G = 2 * tbs * 3
            r0, e0 = o.encode_tb(tbs, 2, 0, G, data)
            r1, e1 = ref.encode_tb(tbs, 2, 0, G, data)
            assert r0 == r1 == 0 and np.array_equal(e0, e1)
            break
    assert found


def test_encode_tb_oracle_errors():
    o = ol.oracle()
    data = np.zeros(1000, np.uint8)
    assert o.encode_tb(1000, 0, 0, 3000, data)[0] == -1      # Qm == 0, sch.c:265-268
    assert o.encode_tb(6152, 2, 0, 30000, data)[0] == -1     # filler bits, sch.c:254-257


def test_encode_decode_round_trip_oracle():
    """encode_tb -> hard-decision LLRs -> decode_tb returns the payload (standard TBS: both CB orders coincide)"""
    o = ol.oracle()
    tbs, Qm, G = 12216, 4, 4 * 9000
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    ret, e = o.encode_tb(tbs, Qm, 0, G, data)
    assert ret == 0
    llr = ((np.unpackbits(e)[:G].astype(np.int16) * 2 - 1) * 40).astype(np.int16)
    d = o.decode_tb(tbs, Qm, 0, llr, 8)
    assert d["ret"] == 0 and np.array_equal(d["data"][:tbs // 8], data)
