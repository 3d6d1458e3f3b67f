"""The reference-side binding (integration/srsran_b200_shim.c): built against the reference's own headers in the
container, exercised on the GPU through the reference's symbol names and struct layouts."""
import ctypes as C
import os
import sys
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import vecgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "integration", "_build", "libsrsran_b200_shim.so")


@pytest.mark.ref
def test_shim_builds_against_reference_headers():
    if not os.path.isdir(os.environ.get("SRSRAN_REF", "/root/reference")):
        pytest.skip("no reference tree")
    import __graft_entry__ as g
    g.build()
    out = subprocess.check_output(["make", "-s", "-C", os.path.join(ROOT, "integration"), "check"]).decode()
    assert "exports all 27" in out


@pytest.mark.ref
def test_pdsch_hook_applies_to_the_reference_and_compiles(tmp_path):
    """integration/apply_b200_patch.py --pdsch on the reference's own pdsch.c: the patched file compiles against the reference
    headers with -DSRSRAN_B200 and calls the three symbol-input hooks; without the define it is the reference's code (the
    hooks vanish)"""
    ref = os.environ.get("SRSRAN_REF", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("no reference tree")
    import __graft_entry__ as g
    g.build()  # oracle/_ref/gen holds the generated config headers the reference sources include
    src = os.path.join(ref, "lib", "src", "phy", "phch", "pdsch.c")
    out_c = str(tmp_path / "pdsch_b200.c")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "apply_b200_patch.py"), "--pdsch", src, out_c])
    patched = open(out_c).read()
    assert patched.count("#ifdef SRSRAN_B200") == 3 and "srsran_b200_dlsch_decode2_symbols(dl_sch, cfg, q->d[codeword_idx]" in patched
    flags = ["-O1", "-std=gnu99", "-mavx2", "-mfma", "-DLV_HAVE_SSE", "-DLV_HAVE_AVX", "-DLV_HAVE_AVX2", "-fPIC", "-w", "-I" + os.path.join(ref, "lib", "include"),
             "-I" + os.path.join(ROOT, "oracle", "_ref", "gen"), "-I" + os.path.join(ref, "lib", "src", "phy", "phch")]
    for define, n_hooks in (("-DSRSRAN_B200", 3), ("-USRSRAN_B200", 0)):
        obj = str(tmp_path / ("pdsch%d.o" % n_hooks))
        subprocess.check_call([os.environ.get("CC", "gcc")] + flags + [define, "-c", out_c, "-o", obj])
        syms = subprocess.check_output(["nm", obj]).decode()
        assert sum(1 for line in syms.splitlines() if " U srsran_b200_" in line) == n_hooks


@pytest.fixture(scope="module")
def shim():
    if not os.path.exists(SHIM):
        pytest.skip("shim not built (needs the reference headers; built in the container and shipped with the repo)")
    return C.CDLL(SHIM)


@pytest.mark.gpu
def test_shim_tdec_symbols(shim):
    """srsran_tdec_init / new_cb / iteration / run_all / get_nof_iterations / free on a real srsran_tdec_t"""
    o = ol.oracle()
    shim.srsran_b200_selftest_sizeof_tdec.restype = C.c_size_t
    h = C.create_string_buffer(shim.srsran_b200_selftest_sizeof_tdec() + 64)
    assert shim.srsran_tdec_init(h, 6144) == 0
    assert shim.srsran_tdec_autoimp_get_subblocks(6144) == 0
    assert shim.srsran_tdec_new_cb(h, 41) == -1
    for K in (40, 6144):
        _, llr = vecgen.make_cb(K, 1.5, 5 + K)
        hard = o.tdec_trace(K, llr, 4)
        out = np.zeros(K // 8, np.uint8)
        assert shim.srsran_tdec_new_cb(h, K) == 0
        for it in range(4):
            shim.srsran_tdec_iteration(h, llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert (out == hard[it]).all()
            assert shim.srsran_tdec_get_nof_iterations(h) == it + 1
        assert shim.srsran_tdec_run_all(h, llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 3, K) == 0
        assert (out == hard[2]).all()
        again = np.zeros(K // 8, np.uint8)
        assert shim.srsran_tdec_get_hard_decision(h, again.ctypes.data_as(C.c_void_p), K) == 0   # north-star alias
        assert (again == hard[2]).all()
        assert shim.srsran_tdec_get_hard_decision(h, again.ctypes.data_as(C.c_void_p), K + 8) == -2
        # 8-bit LLR entry points: the reference's windowed int8 decoder where it has one (K > 800, K % 16 == 0), else widened into
        # the exact int16 engine
        l8 = np.clip(llr // 8, -127, 127).astype(np.int8)
        out8 = np.zeros(K // 8, np.uint8)
        assert shim.srsran_tdec_run_all_8bit(h, l8.ctypes.data_as(C.c_void_p), out8.ctypes.data_as(C.c_void_p), 4, K) == 0
        if o.tdec8_windows(K):
            tr = o.tdec8_trace(K, l8, 4)
            assert (out8 == tr[3]).all()
            assert shim.srsran_tdec_new_cb(h, K) == 0
            for it in range(3):
                shim.srsran_tdec_iteration_8bit(h, l8.ctypes.data_as(C.c_void_p), out8.ctypes.data_as(C.c_void_p))
                assert (out8 == tr[it]).all() and shim.srsran_tdec_get_nof_iterations(h) == it + 1
        else:
            assert (out8 == o.tdec_run_all(K, l8.astype(np.int16), 4)).all()
    shim.srsran_tdec_free(h)


@pytest.mark.gpu
def test_shim_rm_and_decode_tb(shim):
    o = ol.oracle()
    rng = np.random.default_rng(1)
    e = rng.integers(-3000, 3000, 6646).astype(np.int16)
    bo = np.zeros(ol.SOFTBUFFER_SIZE, np.int16); bg = bo.copy()
    o.rm_rx(e, bo, 187, 2)
    assert shim.srsran_rm_turbo_rx_lut(e.ctypes.data_as(C.c_void_p), bg.ctypes.data_as(C.c_void_p), len(e), 187, 2) == 0
    assert (bo == bg).all()
    assert shim.srsran_rm_turbo_rx_lut(e.ctypes.data_as(C.c_void_p), bg.ctypes.data_as(C.c_void_p), len(e), 188, 0) == -2
    # decode_tb through srsran_sch_t / srsran_softbuffer_rx_t / srsran_cbsegm_t, two HARQ transmissions
    tbs, G, Qm = 12216, 19200, 6
    st = None
    Cn = 2
    buf = np.zeros((Cn, ol.SOFTBUFFER_SIZE), np.int16); sbd = np.zeros((Cn, ol.SOFTBUFFER_SIZE // 8), np.uint8)
    cbc = np.zeros(Cn, np.uint8); tbc = np.zeros(1, np.uint8); avg = C.c_float(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for rv in (0, 2):
        _, eb = vecgen.make_tb(tbs, G, Qm, rv, 0.8, 41, scale=100)
        res = o.decode_tb(tbs, Qm, rv, eb, 6, st); st = res["state"]
        data = np.zeros(Cn * 768 + 8, np.uint8)
        ret = shim.srsran_b200_selftest_decode_tb(tbs, Qm, rv, G, p(eb), 6, Cn, p(buf), p(sbd), p(cbc), p(tbc), p(data), C.byref(avg))
        assert ret == res["ret"] and int(tbc[0]) == res["tb_crc"] and (cbc == st["cb_crc"]).all()
        assert np.float32(avg.value) == np.float32(res["avg_iterations"])
        assert (buf == st["buffer_f"]).all() and (sbd == st["sb_data"]).all()
        K1 = res["seg"]["K1"]
        nb = (K1 - 24) // 8 + K1 // 8
        assert (data[:nb] == res["data"][:nb]).all()


@pytest.mark.gpu
def test_shim_encode_tb(shim):
    """srsran_b200_encode_tb through srsran_sch_t / srsran_softbuffer_tx_t / srsran_cbsegm_t: bits at w_offset equal the
    oracle's, every other bit of e_bits is preserved (srsran_bit_copy semantics of encode_tb_off)"""
    o = ol.oracle()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(9)
    for tbs, Qm, rv, G, woff in [(12216, 6, 0, 19200, 0), (12216, 4, 2, 19202, 0), (2984, 2, 1, 4000, 12), (75376, 6, 3, 90000, 24), (40, 2, 0, 121, 3)]:
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        r0, e0 = o.encode_tb(tbs, Qm, rv, G, data)
        nb = Qm * (G // Qm)
        e = rng.integers(0, 256, (G + woff + 7) // 8 + 8, dtype=np.uint8)
        before = np.unpackbits(e)
        ret = shim.srsran_b200_selftest_encode_tb(tbs, Qm, rv, G, p(data), p(e), woff, 64)
        assert ret == r0 == 0
        after = np.unpackbits(e)
        assert np.array_equal(after[woff:woff + nb], np.unpackbits(e0)[:nb])
        assert np.array_equal(after[:woff], before[:woff]) and np.array_equal(after[woff + nb:], before[woff + nb:])
    d = np.zeros(2000, np.uint8); e = np.zeros(8000, np.uint8)
    assert shim.srsran_b200_selftest_encode_tb(6152, 2, 0, 30000, p(d), p(e), 0, 64) == -1   # filler bits
    assert shim.srsran_b200_selftest_encode_tb(12216, 2, 0, 30000, p(d), p(e), 0, 1) == -1   # C > max_cb


@pytest.mark.gpu
def test_shim_ulsch_decode_tb(shim):
    """srsran_b200_ulsch_decode_tb (de-interleave + decode_tb in one device submission) on real reference structs against
    oracle de-interleaver + oracle decode_tb"""
    o = ol.oracle()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    tbs, Qm, nprb, nsymb = 12216, 4, 25, 12
    H = nprb * 12 * nsymb
    rows = H // nsymb
    ri = []
    for m in range(8):
        r, c = rows - 1 - m // 4, (1, 4, 7, 10)[m % 4]
        ri += [r * Qm + c * rows * Qm + k for k in range(Qm)]
    ri = np.array(ri, np.uint32)
    Gb = H * Qm - len(ri)
    Cn = 2
    buf = np.zeros((Cn, ol.SOFTBUFFER_SIZE), np.int16); sbd = np.zeros((Cn, ol.SOFTBUFFER_SIZE // 8), np.uint8)
    cbc = np.zeros(Cn, np.uint8); tbc = np.zeros(1, np.uint8)
    st = None
    for rv in (0, 2):
        _, e = vecgen.make_tb(tbs, Gb, Qm, rv, 1.0, 91, scale=100)
        rng = np.random.default_rng(rv)
        q = rng.integers(-500, 500, H * Qm).astype(np.int16)
        is_ri = np.zeros(H * Qm, bool); is_ri[ri] = True
        j, i, k = np.meshgrid(np.arange(rows), np.arange(nsymb), np.arange(Qm), indexing="ij")
        pos = (j * Qm + i * rows * Qm + k).reshape(-1)
        pos = pos[~is_ri[pos]]
        q[pos] = e
        g_ref = o.ulsch_deinterleave(q, Qm, H, nsymb, list(ri))
        res = o.decode_tb(tbs, Qm, rv, g_ref[:Gb], 6, st); st = res["state"]
        data = np.zeros(Cn * 768 + 8, np.uint8)
        g = np.zeros(16, np.int16)
        ret = shim.srsran_b200_selftest_ulsch_decode_tb(tbs, Qm, rv, p(q), H, nsymb, p(ri), len(ri), 0, Gb, p(g), 16, 6, Cn, p(buf), p(sbd), p(cbc), p(tbc),
                                                        p(data))
        assert ret == res["ret"] and int(tbc[0]) == res["tb_crc"] and (cbc == st["cb_crc"]).all()
        assert (buf == st["buffer_f"]).all() and (g == g_ref[:16]).all()
        if ret == 0:
            assert (data[:tbs // 8] == res["data"][:tbs // 8]).all()
    assert ret == 0


@pytest.mark.gpu
def test_shim_thread_engines_are_destroyed_at_thread_exit(shim):
    """the per-thread engines must not outlive their thread (streams, pinned arenas, device tables). 24 short-lived
    threads each decode one K=6144 block through srsran_tdec_run_all; the device memory in use afterwards must be back to where it
    was after the first few (a leak of one engine per thread would pile up tens of MB each)."""
    import threading
    import torch
    o = ol.oracle()
    K = 6144
    _, llr = vecgen.make_cb(K, 1.5, 77)
    ref = o.tdec_run_all(K, llr, 4)
    shim.srsran_b200_selftest_sizeof_tdec.restype = C.c_size_t
    sz = shim.srsran_b200_selftest_sizeof_tdec() + 64
    errs = []

    def work():
        try:
            h = C.create_string_buffer(sz)
            out = np.zeros(K // 8, np.uint8)
            assert shim.srsran_tdec_init(h, K) == 0
            assert shim.srsran_tdec_run_all(h, llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 4, K) == 0
            assert (out == ref).all()
            shim.srsran_tdec_free(h)
        except BaseException as ex:  # noqa: BLE001
            errs.append(ex)

    def run(n):
        for _ in range(n):
            t = threading.Thread(target=work)
            t.start()
            t.join()

    run(4)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    run(20)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert not errs, errs[0]
    assert free0 - free1 < 32 * 2 ** 20, "device memory grew by %.1f MB over 20 thread lifetimes" % ((free0 - free1) / 2 ** 20)
    assert shim.srsran_b200_nof_devices() >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("tbs,nre,mod,eb,layers", [(75376, 14400, 3, 12.5, 1), (12216, 4803, 2, 6.5, 1), (6120, 7200, 1, 2.5, 1), (36696, 7200, 3, 12.0, 2)])
def test_pdsch_codeword_from_symbols(shim, tbs, nre, mod, eb, layers):
    """what the patched srsran_pdsch_codeword_decode calls (pdsch.c:693-740 under -DSRSRAN_B200): srsran_b200_dlsch_decode2_symbols on
    a real srsran_pdsch_cfg_t - equalised symbols in, payload out - against the oracle chain soft demodulation -> descrambling with
    the PDSCH seed (sequences.c:62-65) -> decode_tb; one codeword on one and on two layers (Qm x Nl for the rate de-matcher)"""
    o = ol.oracle()
    Qm = (1, 2, 4, 6, 8)[mod]
    G = nre * Qm
    rnti, cw, nslot, cell = 0x4601, 0, 2 * 7, 301
    c_init = (rnti << 14) + (cw << 13) + ((nslot // 2) << 9) + cell
    assert shim.srsran_b200_pdsch_c_init(rnti, cw, nslot, cell) == c_init
    rng = np.random.default_rng(tbs + mod + layers)
    payload, e_clean = vecgen.make_tb(tbs, G, Qm * layers, 0, 60.0, 15 + mod, scale=8)
    scr = o.sequence_apply_s(np.ones(G, np.int16), c_init) < 0
    bits = (e_clean > 0).astype(np.uint8) ^ scr.astype(np.uint8)
    b = 1.0 - 2.0 * bits.astype(np.float64).reshape(-1, Qm)
    if mod == 1:
        sym = (b[:, 0] + 1j * b[:, 1]) / np.sqrt(2.0)
    elif mod == 2:
        sym = (b[:, 0] * (2.0 - b[:, 2]) + 1j * b[:, 1] * (2.0 - b[:, 3])) / np.sqrt(10.0)
    else:
        sym = (b[:, 0] * (4.0 - b[:, 2] * (2.0 - b[:, 4])) + 1j * b[:, 1] * (4.0 - b[:, 3] * (2.0 - b[:, 5]))) / np.sqrt(42.0)
    sigma = vecgen.sigma_for(eb, tbs / float(G)) / np.sqrt(float(Qm))
    sym = (sym + sigma * (rng.standard_normal(nre) + 1j * rng.standard_normal(nre))).astype(np.complex64)
    llr = o.sequence_apply_s(o.demod_soft_demodulate_s(mod, sym), c_init)
    res = o.decode_tb(tbs, Qm * layers, 0, llr, 8)
    Cn = res["seg"]["C"]
    buf = np.zeros((Cn, ol.SOFTBUFFER_SIZE), np.int16); sbd = np.zeros((Cn, ol.SOFTBUFFER_SIZE // 8), np.uint8)
    cbc = np.zeros(Cn, np.uint8); tbc = np.zeros(1, np.uint8); avg = C.c_float(0)
    data = np.zeros(tbs // 8 + 64, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    shim.srsran_b200_selftest_dlsch_symbols.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint16, C.c_int,
                                                        C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_void_p]
    ret = shim.srsran_b200_selftest_dlsch_symbols(tbs, mod, 0, nre, 1, layers, p(sym), rnti, cw, nslot, cell, 8, Cn, p(buf), p(sbd), p(cbc), p(tbc), p(data),
                                                  C.cast(C.byref(avg), C.c_void_p))
    assert ret == res["ret"] == 0 and int(tbc[0]) == res["tb_crc"] and (cbc == res["state"]["cb_crc"][:Cn]).all()
    assert np.float32(avg.value) == np.float32(res["avg_iterations"])
    assert (buf == res["state"]["buffer_f"][:Cn]).all()
    assert (data[:tbs // 8] == res["data"][:tbs // 8]).all() and (data[:tbs // 8] == payload[:tbs // 8]).all()
