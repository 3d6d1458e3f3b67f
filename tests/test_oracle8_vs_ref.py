"""The 8-bit LLR mode: oracle/turbo_oracle8.c - the reference's windowed saturating int8 decoders restated
in natural order - PINNED against the compiled reference (oracle/_ref, AVX2 build): srsran_tdec_iteration_8bit in AUTO mode,
srsran_rm_turbo_rx_lut_8bit, and the literal sch.c loop with q->llr_is_8bit."""
import numpy as np
import pytest

import oracle_lib as ol
import vecgen

ref = ol.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built")


def llr8(K, eb, seed, scale=16):
    _, l16 = vecgen.make_cb(K, eb, seed, scale=scale)
    return np.clip(l16, -127, 127).astype(np.int8)


def test_window_counts_match_autoimp():
    o = ol.oracle()
    for idx in range(188):
        K = o.cbsize(idx)
        nw = ref.lib.srsran_tdec_autoimp_get_subblocks_8bit(K) if hasattr(ref.lib, "srsran_tdec_autoimp_get_subblocks_8bit") else None
        exp = 32 if (K % 32 == 0 and K > 2048) else (16 if (K % 16 == 0 and K > 800) else 0)
        assert o.tdec8_windows(K) == exp
        if nw is not None:
            assert (nw if nw in (16, 32) else 0) == exp


# every size class: 16 windows with and without the K % 32 == 16 tail of srsran_vec_sub_bbb, 32 windows small / config 3 / largest
@pytest.mark.parametrize("K", [816, 832, 1008, 1024, 1056, 2048, 2112, 3008, 5824, 6016, 6144])
@pytest.mark.parametrize("eb,scale", [(1.5, 16), (0.5, 40), (4.0, 8), (8.0, 60)])
def test_tdec8_soft_outputs_and_decisions_vs_reference(K, eb, scale):
    """ext1 / ext2 / app1 after every half-iteration, element by element, and the hard decisions - including heavy saturation
    (scale 40-60 puts most input LLRs at +-127)"""
    o = ol.oracle()
    l8 = llr8(K, eb, 1000 + K, scale)
    n_it = 6
    out_o, d_o = o.tdec8_trace(K, l8, n_it, dump=True)
    out_r, d_r = ref.tdec8_trace(K, l8, n_it, dump=True)
    for it in range(n_it):
        names = ("ext1", "ext2", "app1")
        for a in range(3):
            if it == 0 and a > 0:
                continue   # ext2 / app1 are not written by the first half-iteration
            assert np.array_equal(d_o[it, a], d_r[it, a]), (K, it, names[a], np.flatnonzero(d_o[it, a] != d_r[it, a])[:8])
        assert np.array_equal(out_o[it], out_r[it]), (K, it)


def test_tdec8_random_inputs_full_range():
    """uniformly random int8 inputs (no code structure): the arithmetic itself, every saturation corner"""
    o = ol.oracle()
    rng = np.random.default_rng(5)
    for K in (1008, 2560, 6144):
        l8 = rng.integers(-128, 128, 3 * K + 12).astype(np.int8)
        out_o, d_o = o.tdec8_trace(K, l8, 4, dump=True)
        out_r, d_r = ref.tdec8_trace(K, l8, 4, dump=True)
        assert np.array_equal(d_o[3], d_r[3]) and np.array_equal(out_o, out_r)


def _sb_to_natural(buf, K, nw):
    """srsran_rm_turbo_rx_lut_8bit's sub-block layout (rm_turbo.c:260-273: planes at 0, K+32, 2(K+32), termination at 3(K+32), each
    plane window-interleaved) back to the natural s p p' order"""
    S = K // nw
    n = np.arange(K)
    idx = (n % S) * nw + n // S
    nat = np.zeros(3 * K + 12, buf.dtype)
    nat[0:3 * K:3] = buf[idx]
    nat[1:3 * K:3] = buf[K + 32 + idx]
    nat[2:3 * K:3] = buf[2 * (K + 32) + idx]
    nat[3 * K:] = buf[3 * (K + 32):3 * (K + 32) + 12]
    return nat


@pytest.mark.parametrize("K", [1008, 2048, 5824, 6144])
def test_rm_rx8_vs_reference(K):
    """output[T[i mod L]] += input[i] in wrapping int8: puncturing, exact fit, repetition (the SSE wrap path), accumulation over
    two redundancy versions; the reference result is brought from its sub-block layout to natural order"""
    o = ol.oracle()
    idx = o.cbindex(K)
    nw = o.tdec8_windows(K)
    rng = np.random.default_rng(K)
    L = 3 * K + 12
    for E in (L // 2 + 3, L, L + 16 * 7 + 5, 2 * L + 999):
        b_o = np.zeros(ol.SOFTBUFFER_SIZE, np.int8)
        b_r = np.zeros(ol.SOFTBUFFER_SIZE + 256, np.int8)
        for rv in (0, 2):
            e = rng.integers(-128, 128, E).astype(np.int8)
            o.rm_rx8(e, b_o, idx, rv)
            ref.rm_rx8(e, b_r, idx, rv)
        assert np.array_equal(b_o[:L], _sb_to_natural(b_r, K, nw)), (K, E)


@pytest.mark.parametrize("tbs,Qm,G", [(75376, 6, 86400), (149776, 6, 12 * 14400), (36696, 6, 6 * 8000), (12960, 4, 4 * 5000)])
def test_decode_tb8_loop_vs_literal_sch(tbs, Qm, G):
    """the restated loop with int8 soft buffers == srsran_dlsch_decode2 with q->llr_is_8bit (the SAME decoder on both sides, so
    the half-iteration counts are pinned too): return code, bytes, flags, average iterations over rv 0, 2"""
    o = ol.oracle()
    _, seg = o.cbsegm(tbs)
    assert o.tdec8_windows(seg["K1"])
    Qe = Qm * (2 if seg["C"] > 16 else 1)
    h = ref.dlsch_rx_new_guru(32)
    st = None
    try:
        for rv, eb in ((0, 1.0), (2, 1.0)):
            _, e16 = vecgen.make_tb(tbs, G, Qe, rv, eb, 31 + tbs, scale=12)
            e8 = np.clip(e16, -127, 127).astype(np.int8)
            a = o.decode_tb8(tbs, Qe, rv, e8, 8, st)
            st = a["state"]
            if seg["C"] > 16:
                b = None   # (ref_dlsch_decode8 binds codeword 0 / one layer: the two-layer TB goes through the loop check below only)
            else:
                b = ref.dlsch_decode8(h, tbs, Qm, rv, e8, 8)
            if b is not None:
                assert a["ret"] == b["ret"] and a["tb_crc"] == b["tb_crc"]
                assert np.array_equal(st["cb_crc"][:seg["C"]], b["cb_crc"][:seg["C"]])
                assert np.float32(a["avg_iterations"]) == np.float32(ref.lib.ref_last_avg_iterations())
                if a["ret"] == 0:
                    assert np.array_equal(a["data"][:tbs // 8], b["data"][:tbs // 8])
    finally:
        ref.dlsch_rx_free(h)
