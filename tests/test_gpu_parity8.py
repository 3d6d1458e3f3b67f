"""8-bit LLR mode on the GPU: the CUDA path through the C ABI against oracle/turbo_oracle8.c - itself pinned to
the compiled reference's srsran_tdec_iteration_8bit / srsran_rm_turbo_rx_lut_8bit / llr_is_8bit loop (tests/test_oracle8_vs_ref.py).
Integer work: bit-exact hard bits, half-iteration counts, CRC verdicts, soft buffers."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import vecgen

pytestmark = pytest.mark.gpu


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


def llr8_batch(K, n, eb, seed, scale=16):
    _, l16 = vecgen.make_cb_batch(K, n, eb, seed, scale=scale)
    return np.clip(l16, -127, 127).astype(np.int8)


def test_windows_table(sb, o):
    for idx in range(188):
        K = o.cbsize(idx)
        assert sb.tdec8_windows(K) == o.tdec8_windows(K)
    assert sb.tdec8_windows(6145) == 0


@pytest.mark.parametrize("K", [816, 832, 1008, 1024, 1056, 2048, 2112, 3008, 5824, 6016, 6144])
def test_hard_bits_every_half_iteration(sb, eng, o, K):
    """every window geometry (16 / 32 windows, window lengths that are / are not multiples of 8, the K % 32 == 16 wrap tail of the
    8-bit subtraction), half-iterations 1..6 without early stop, moderate and heavily saturated inputs"""
    for eb, scale in ((1.5, 16), (0.5, 40), (6.0, 60)):
        l8 = llr8_batch(K, 3, eb, 3000 + K, scale)
        traces = [o.tdec8_trace(K, l8[c], 6) for c in range(3)]
        for it in range(1, 7):
            out, noi, _ = eng.tdec_batch8(K, l8, it, early_stop=False, crc_kind=sb.CRC_NONE)
            assert (noi == it).all()
            for c in range(3):
                assert (out[c] == traces[c][it - 1]).all(), (K, eb, scale, it, c)


def test_random_full_range_inputs(sb, eng, o):
    rng = np.random.default_rng(8)
    for K in (1008, 2560, 6144):
        l8 = rng.integers(-128, 128, (5, 3 * K + 12)).astype(np.int8)
        out, _, _ = eng.tdec_batch8(K, l8, 5, early_stop=False, crc_kind=sb.CRC_NONE)
        for c in range(5):
            assert (out[c] == o.tdec8_trace(K, l8[c], 5)[4]).all(), (K, c)


@pytest.mark.parametrize("K,n", [(6144, 1), (6144, 2), (6144, 37), (1024, 1), (1024, 3), (1024, 4), (1024, 9), (2048, 130)])
def test_batch_early_stop_odd_counts(sb, eng, o, K, n):
    """CRC24B early stop, iteration counts, verdicts; unit packing with empty halves / pairs"""
    l8 = llr8_batch(K, n, 1.8 if K > 2000 else 2.5, 77 + n, 14)
    _, oo, on, ook = o.tdec8_batch(K, l8, 8, True, nthreads=4)
    out, noi, ok = eng.tdec_batch8(K, l8, 8, early_stop=True)
    assert (noi == on).all() and (ok == ook).all() and (out == oo).all()
    assert ok.mean() > 0.5


def test_every_8bit_block_size_in_one_submission(sb, eng, o):
    """all LTE block sizes the reference decodes in 8-bit arithmetic (K > 800 and K % 16 == 0: 16 windows up to 2048, 32 above where
    K % 32 == 0), three blocks each in ONE mixed submission (a warp unit per size and pair; window lengths 51..192, the short last
    recompute block, the wrap-around tail of K % 32 != 0), every block against the 8-bit oracle"""
    rng = np.random.default_rng(808 + int(os.environ.get("SRSB200_FUZZ_SEED", "0")))
    sizes = [o.cbsize(i) for i in range(188)]
    sizes = [k for k in sizes if sb.tdec8_windows(k)]
    assert len(sizes) > 90 and min(sizes) == 816 and max(sizes) == 6144
    Ks, llrs = [], []
    for K in sizes:
        for j in range(3):
            _, l = vecgen.make_cb(K, float(rng.choice([0.5, 1.5, 2.5, 4.0])), int(rng.integers(1 << 30)), scale=int(rng.choice([8, 12, 20])))
            Ks.append(K)
            llrs.append(np.clip(l, -127, 127).astype(np.int8))
    outs, noi, ok = eng.tdec_batch8_mixed(np.array(Ks, np.uint32), llrs, 8)
    for i, K in enumerate(Ks):
        _, oo, on, ook = o.tdec8_batch(K, llrs[i][None, :], 8, True)
        assert on[0] == noi[i] and ook[0] == ok[i] and np.array_equal(oo[0], outs[i]), (K, i % 3)


def test_rm_rx8(sb, eng, o):
    rng = np.random.default_rng(3)
    for K in (1008, 6144):
        idx = o.cbindex(K)
        L = 3 * K + 12
        for E in (L // 2 + 3, L, 2 * L + 999):
            b_o = rng.integers(-128, 128, ol.SOFTBUFFER_SIZE).astype(np.int8)
            b_g = b_o.copy()
            for rv in (0, 3):
                e = rng.integers(-128, 128, E).astype(np.int8)
                o.rm_rx8(e, b_o, idx, rv)
                assert eng.rm_turbo_rx_lut8(e, b_g, idx, rv) == 0
            assert np.array_equal(b_o[:L], b_g[:L])


@pytest.mark.parametrize("tbs,G,Qe,eb", [(75376, 86400, 6, 1.0), (149776, 12 * 14400, 12, 1.0), (36696, 6 * 8000, 6, 0.3), (12960, 4 * 5000, 4, 0.5)])
def test_decode_tb8_harq(sb, eng, o, tbs, G, Qe, eb):
    """decode_tb with llr_is_8bit over rv 0, 2, 3, 1: return code, bytes, per-block iteration counts, flags, int8 soft buffers"""
    tb = sb.TransportBlock(tbs)
    st = None
    C_ = tb.seg["C"]
    rets = []
    for rv in (0, 2, 3, 1):
        _, e16 = vecgen.make_tb(tbs, G, Qe, rv, eb, 131 + tbs, scale=12)
        e8 = np.clip(e16, -127, 127).astype(np.int8)
        a = o.decode_tb8(tbs, Qe, rv, e8, 8, st)
        st = a["state"]
        tb.data[:] = 0
        assert eng.decode_tb(tb, Qe, rv, e8, 8, llr8=True) == a["ret"]
        rets.append(a["ret"])
        assert (tb.cb_noi[:C_] == a["cb_noi"][:C_]).all()
        assert (tb.cb_crc[:C_] == st["cb_crc"][:C_]).all() and tb.tb_crc[0] == a["tb_crc"]
        assert np.float32(tb.avg_iterations) == np.float32(a["avg_iterations"])
        K1 = a["seg"]["K1"]
        nbytes = (C_ - 1) * ((K1 - 24) // 8) + K1 // 8 if C_ > 1 else K1 // 8
        assert (tb.data[:nbytes] == a["data"][:nbytes]).all()
        Lb = 3 * K1 + 12
        assert np.array_equal(tb.buffer_b[:C_, :Lb], st["buffer_b"][:C_, :Lb])
        assert (tb.sb_data[:C_] == st["sb_data"][:C_]).all()
    assert 0 in rets


def test_decode_tb8_refuses_small_blocks_and_mixed_flags(sb, eng):
    tb = sb.TransportBlock(6200)   # C = 2, K = 3136? no: 6200+24 -> two blocks of 3136: has an 8-bit decoder
    small = sb.TransportBlock(1000)  # one block of K = 1024 ... has 16 windows; 40 bits -> K = 64: none
    tiny = sb.TransportBlock(40)
    assert eng.decode_tb(tiny, 2, 0, np.zeros(300, np.int8), 8, llr8=True) == -2
    assert eng.decode_tb_batch([(tb, 4, 0, np.zeros(9000, np.int8))], 8, llr8=True) == 0
    del small
