"""ctypes bindings for the test oracle (oracle/libturbo_oracle.so) and, when it has been built in the container,
the compiled reference (oracle/_ref/libsrsran_ref.so). TEST INFRASTRUCTURE: imported only by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libturbo_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsrsran_ref.so")

CRC24A = 0x1864CFB
CRC24B = 0x1800063
SOFTBUFFER_SIZE = 18600

i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("turbo_oracle.c", "turbo_oracle8.c", "turbo_oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


def build_ref():
    """Compile the reference sources in place (container only: needs /root/reference)."""
    if os.path.isdir(os.environ.get("SRSRAN_REF", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


class _Lib:
    """Common surface of the two libraries; `p` is the symbol prefix ('orc_' or 'ref_')."""

    def __init__(self, path, p):
        self.lib = C.CDLL(path)
        self.p = p
        L = self.lib
        f = lambda n: getattr(L, p + n)
        self._cbsize = f("cbsize"); self._cbsize.argtypes = [C.c_uint32]; self._cbsize.restype = C.c_int
        self._cbindex = f("cbindex"); self._cbindex.argtypes = [C.c_uint32]; self._cbindex.restype = C.c_int
        self._cbsegm = f("cbsegm"); self._cbsegm.argtypes = [C.c_uint32, u32p]; self._cbsegm.restype = C.c_int
        self._qpp = f("qpp"); self._qpp.argtypes = [C.c_uint32, u16p, u16p]; self._qpp.restype = C.c_int
        self._crc_bytes = f("crc_bytes"); self._crc_bytes.argtypes = [C.c_uint32, C.c_int, u8p, C.c_int]; self._crc_bytes.restype = C.c_uint32
        self._crc_bits = f("crc_bits"); self._crc_bits.argtypes = [C.c_uint32, C.c_int, u8p, C.c_int]; self._crc_bits.restype = C.c_uint32
        self._enc = f("tcod_encode"); self._enc.argtypes = [u8p, u8p, C.c_uint32]; self._enc.restype = C.c_int
        self._rm_tx = f("rm_tx"); self._rm_tx.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_uint32]; self._rm_tx.restype = C.c_int
        self._rm_table = f("rm_table"); self._rm_table.argtypes = [C.c_uint32, C.c_uint32, u16p]; self._rm_table.restype = C.c_int
        self._rm_rx = f("rm_rx"); self._rm_rx.argtypes = [i16p, i16p, C.c_uint32, C.c_uint32, C.c_uint32]; self._rm_rx.restype = C.c_int
        self._map = f("map_gen"); self._map.argtypes = [C.c_uint32, i16p, C.c_void_p, i16p, i16p]; self._map.restype = C.c_int
        self._dec_tb = f("decode_tb")
        self._dec_tb.argtypes = [C.c_uint32] * 4 + [i16p, C.c_uint32, i16p, u8p, u8p, u8p, u8p, u32p, C.POINTER(C.c_float)]
        self._dec_tb.restype = C.c_int
        self._demod = f("demod_soft_demodulate_s")
        self._demod.argtypes = [C.c_int, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), i16p, C.c_int]; self._demod.restype = C.c_int
        self._seq = f("sequence_apply_s"); self._seq.argtypes = [i16p, i16p, C.c_uint32, C.c_uint32]; self._seq.restype = None
        self._enc_tb = f("encode_tb")
        self._enc_tb.argtypes = [C.c_uint32] * 4 + [u8p, u8p]
        self._enc_tb.restype = C.c_int
        if p == "ref_":
            L.ref_init()
            self._trace = L.ref_tdec_trace; self._trace.argtypes = [C.c_int, C.c_uint32, i16p, C.c_uint32, u8p, C.c_void_p]
            self._run_all = L.ref_tdec_run_all; self._run_all.argtypes = [C.c_int, C.c_uint32, i16p, C.c_uint32, u8p]
            self._batch = L.ref_tdec_batch
            self._batch.argtypes = [C.c_int, C.c_uint32, i16p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, u8p, u8p, u8p]
            self._rm_rx_auto = L.ref_rm_rx_auto; self._rm_rx_auto.argtypes = [i16p, i16p, C.c_uint32, C.c_uint32, C.c_uint32]
            # the literal sch.c entry points (srsran_dlsch_encode2 / srsran_dlsch_decode2 / ulsch_deinterleave)
            L.ref_dlsch_encode.argtypes = [C.c_uint32] * 4 + [u8p, u8p]; L.ref_dlsch_encode.restype = C.c_int
            L.ref_dlsch_rx_new.restype = C.c_void_p
            L.ref_dlsch_rx_new_guru.argtypes = [C.c_uint32]; L.ref_dlsch_rx_new_guru.restype = C.c_void_p
            L.ref_dlsch_rx_free.argtypes = [C.c_void_p]
            L.ref_dlsch_rx_reset.argtypes = [C.c_void_p, C.c_uint32]
            L.ref_dlsch_rx_max_cb.argtypes = [C.c_void_p]; L.ref_dlsch_rx_max_cb.restype = C.c_uint32
            L.ref_dlsch_decode.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [i16p, C.c_uint32, u8p, u8p, u8p, C.POINTER(C.c_float)]
            L.ref_dlsch_decode.restype = C.c_int
            L.ref_dlsch_decode_cw.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [i16p, C.c_uint32, u8p, u8p, u8p, C.POINTER(C.c_float)] + [C.c_uint32] * 3
            L.ref_dlsch_decode_cw.restype = C.c_int
            L.ref_dlsch_encode_retx_null.argtypes = [C.c_uint32] * 4 + [u8p, u8p]; L.ref_dlsch_encode_retx_null.restype = C.c_int
            L.ref_dlsch_encode_cw.argtypes = [C.c_uint32] * 4 + [u8p, u8p] + [C.c_uint32] * 3; L.ref_dlsch_encode_cw.restype = C.c_int
            L.ref_ulsch_deinterleave.argtypes = [i16p, C.c_uint32, C.c_uint32, C.c_uint32, i16p, u32p, C.c_uint32]
            L.ref_dlsch_decode8.argtypes = [C.c_void_p] + [C.c_uint32] * 4 + [i8p, C.c_uint32, u8p, u8p, u8p]
            L.ref_last_avg_iterations.restype = C.c_float
            L.ref_tdec8_trace.argtypes = [C.c_uint32, i8p, C.c_uint32, u8p, C.c_void_p]; L.ref_tdec8_trace.restype = C.c_int
            L.ref_rm_rx8.argtypes = [i8p, i8p, C.c_uint32, C.c_uint32, C.c_uint32]; L.ref_rm_rx8.restype = C.c_int
            L.ref_dlsch_decode8.restype = C.c_int
            L.ref_ulsch_encode.argtypes = [C.c_uint32] * 7 + [u8p, u8p]; L.ref_ulsch_encode.restype = C.c_int
            L.ref_ulsch_decode.argtypes = [C.c_void_p] + [C.c_uint32] * 6 + [i16p, C.c_uint32, u8p, u8p, u8p, C.POINTER(C.c_float), u8p]
            L.ref_ulsch_decode.restype = C.c_int
        else:
            self._trace = L.orc_tdec_trace; self._trace.argtypes = [C.c_uint32, i16p, C.c_uint32, u8p, C.c_void_p]
            self._run_all = L.orc_tdec_run_all; self._run_all.argtypes = [C.c_uint32, i16p, C.c_uint32, u8p]
            L.orc_tdec8_windows.argtypes = [C.c_uint32]; L.orc_tdec8_windows.restype = C.c_uint32
            L.orc_tdec8_trace.argtypes = [C.c_uint32, i8p, C.c_uint32, u8p, C.c_void_p]; L.orc_tdec8_trace.restype = C.c_int
            L.orc_tdec8_batch.argtypes = [C.c_uint32, i8p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, u8p, u8p, u8p]; L.orc_tdec8_batch.restype = C.c_double
            L.orc_rm_rx8.argtypes = [i8p, i8p, C.c_uint32, C.c_uint32, C.c_uint32]; L.orc_rm_rx8.restype = C.c_int
            L.orc_decode_tb8.argtypes = [C.c_uint32] * 4 + [i8p, C.c_uint32, i8p, u8p, u8p, u8p, u8p, u32p, C.POINTER(C.c_float)]
            L.orc_decode_tb8.restype = C.c_int
            self._batch = L.orc_tdec_batch
            self._batch.argtypes = [C.c_uint32, i16p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, u8p, u8p, u8p]
            L.orc_ulsch_deinterleave.argtypes = [i16p, C.c_uint32, C.c_uint32, C.c_uint32, i16p, u32p, C.c_uint32]
            L.orc_ulsch_deinterleave.restype = C.c_int
        self._trace.restype = C.c_int
        self._run_all.restype = C.c_int
        self._batch.restype = C.c_double

    # ---- tables
    def cbsize(self, idx):
        return self._cbsize(idx)

    def cbindex(self, K):
        return self._cbindex(K)

    def cbsegm(self, tbs):
        out = np.zeros(12, np.uint32)
        ret = self._cbsegm(tbs, out)
        keys = ["F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs", "L_tb", "L_cb", "Z"]
        return ret, dict(zip(keys, (int(v) for v in out)))

    def qpp(self, K):
        f = np.zeros(K, np.uint16); r = np.zeros(K, np.uint16)
        assert self._qpp(K, f, r) == 0
        return f, r

    def crc_bytes(self, poly, data, nbits, order=24):
        return self._crc_bytes(poly, order, np.ascontiguousarray(data, np.uint8), nbits)

    def crc_bits(self, poly, bits, order=24):
        bits = np.ascontiguousarray(bits, np.uint8)
        return self._crc_bits(poly, order, bits, len(bits))

    # ---- encoder chain
    def encode(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        K = len(bits)
        out = np.zeros(3 * K + 12, np.uint8)
        assert self._enc(bits, out, K) == 0
        return out

    def rm_tx(self, coded, K, E, rv):
        e = np.zeros(E, np.uint8)
        assert self._rm_tx(np.ascontiguousarray(coded, np.uint8), K, e, E, rv) == 0
        return e

    def rm_table(self, cb_idx, rv):
        K = self.cbsize(cb_idx)
        t = np.zeros(3 * K + 12, np.uint16)
        ret = self._rm_table(cb_idx, rv, t)
        assert ret == 0, ret
        return t

    def rm_rx(self, e, buf, cb_idx, rv):
        """accumulates into buf (int16, at least 3K+12 (+slack for the reference's SIMD stores))"""
        e = np.ascontiguousarray(e, np.int16)
        return self._rm_rx(e, buf, len(e), cb_idx, rv)

    # ---- decoder
    def map_gen(self, K, inp, app, par):
        out = np.zeros(K, np.int16)
        inp = np.ascontiguousarray(inp, np.int16); par = np.ascontiguousarray(par, np.int16)
        if app is not None:
            app = np.ascontiguousarray(app, np.int16)
            appp = app.ctypes.data_as(C.c_void_p)
        else:
            appp = None
        assert self._map(K, inp, appp, par, out) == 0
        return out

    def tdec_trace(self, K, llr, nof_iter, dump=False, impl=1):
        llr = np.ascontiguousarray(llr, np.int16)
        out = np.zeros((nof_iter, K // 8), np.uint8)
        d = np.zeros((nof_iter, 3, K), np.int16) if dump else None
        dp = d.ctypes.data_as(C.c_void_p) if dump else None
        if self.p == "ref_":
            ret = self._trace(impl, K, llr, nof_iter, out, dp)
        else:
            ret = self._trace(K, llr, nof_iter, out, dp)
        assert ret == 0, ret
        return (out, d) if dump else out

    def tdec_run_all(self, K, llr, nof_iter, impl=1):
        llr = np.ascontiguousarray(llr, np.int16)
        out = np.zeros(K // 8, np.uint8)
        ret = self._run_all(impl, K, llr, nof_iter, out) if self.p == "ref_" else self._run_all(K, llr, nof_iter, out)
        assert ret == 0, ret
        return out

    def tdec_batch(self, K, llr, max_iter, early_stop, nthreads=1, impl=1, pin=0):
        """llr [n, 3K+12] int16 -> (seconds, out[n,K/8], noi[n], crc_ok[n])"""
        llr = np.ascontiguousarray(llr, np.int16)
        n = llr.shape[0]
        out = np.zeros((n, K // 8), np.uint8); noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
        if self.p == "ref_":
            secs = self._batch(impl, K, llr, n, max_iter, int(early_stop), nthreads, pin, out, noi, ok)
        else:
            secs = self._batch(K, llr, n, max_iter, int(early_stop), nthreads, out, noi, ok)
        return secs, out, noi, ok

    # ---- transport block
    def decode_tb(self, tbs, Qm, rv, e_bits, max_iterations, state=None, nof_e_bits=None):
        """state = dict(buffer_f[C,18600] i16, sb_data[C,18600/8] u8, cb_crc[C] u8) persists across HARQ transmissions"""
        _, seg = self.cbsegm(tbs)
        Cn = max(seg["C"], 1)
        if state is None:
            state = new_tb_state(Cn)
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        G = len(e_bits) if nof_e_bits is None else nof_e_bits
        data = np.zeros(Cn * 768 + 8, np.uint8)
        noi = np.zeros(Cn, np.uint32)
        tb_crc = np.zeros(1, np.uint8)
        avg = C.c_float(0)
        ret = self._dec_tb(tbs, Qm, rv, G, e_bits, max_iterations, state["buffer_f"], state["sb_data"], state["cb_crc"],
                           tb_crc, data, noi, C.byref(avg))
        return dict(ret=ret, data=data, cb_noi=noi, tb_crc=int(tb_crc[0]), avg_iterations=avg.value, state=state, seg=seg)

    # ---- 8-bit LLR mode
    def tdec8_windows(self, K):
        return int(self.lib.orc_tdec8_windows(K))

    def tdec8_trace(self, K, llr8, nof_iter, dump=False):
        """hard decisions after every half-iteration of the 8-bit window decoder; dump -> (out, [it][3][K] ext1, ext2, app1)"""
        llr8 = np.ascontiguousarray(llr8, np.int8)
        out = np.zeros((nof_iter, K // 8), np.uint8)
        d = np.zeros((nof_iter, 3, K), np.int8) if dump else None
        dp = d.ctypes.data_as(C.c_void_p) if dump else None
        ret = getattr(self.lib, self.p + "tdec8_trace")(K, llr8, nof_iter, out, dp)
        assert ret >= 0, ret
        return (out, d) if dump else out

    def tdec8_batch(self, K, llr8, max_iter, early_stop, nthreads=1):
        llr8 = np.ascontiguousarray(llr8, np.int8)
        n = llr8.shape[0]
        out = np.zeros((n, K // 8), np.uint8); noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
        secs = self.lib.orc_tdec8_batch(K, llr8, n, max_iter, int(early_stop), nthreads, out, noi, ok)
        assert secs >= 0
        return secs, out, noi, ok

    def rm_rx8(self, e8, buf8, cb_idx, rv):
        """in-place accumulate into buf8 (oracle: natural layout; compiled reference: its sub-block layout)"""
        e8 = np.ascontiguousarray(e8, np.int8)
        assert getattr(self.lib, self.p + "rm_rx8")(e8, buf8, len(e8), cb_idx, rv) == 0
        return buf8

    def decode_tb8(self, tbs, Qm, rv, e_bits8, max_iterations, state=None):
        """the decode_tb loop with q->llr_is_8bit; state = dict(buffer_b[C,18600] i8, sb_data, cb_crc)"""
        _, seg = self.cbsegm(tbs)
        Cn = max(seg["C"], 1)
        if state is None:
            state = dict(buffer_b=np.zeros((Cn, SOFTBUFFER_SIZE), np.int8), sb_data=np.zeros((Cn, SOFTBUFFER_SIZE // 8), np.uint8),
                         cb_crc=np.zeros(Cn, np.uint8))
        e = np.ascontiguousarray(e_bits8, np.int8)
        data = np.zeros(Cn * 768 + 8, np.uint8); noi = np.zeros(Cn, np.uint32); tb_crc = np.zeros(1, np.uint8); avg = C.c_float(0)
        ret = self.lib.orc_decode_tb8(tbs, Qm, rv, len(e), e, max_iterations, state["buffer_b"], state["sb_data"], state["cb_crc"], tb_crc, data, noi,
                                      C.byref(avg))
        return dict(ret=ret, data=data, cb_noi=noi, tb_crc=int(tb_crc[0]), avg_iterations=avg.value, state=state, seg=seg)

    def demod_soft_demodulate_s(self, mod, symbols):
        """mod 0..4 = BPSK, QPSK, 16QAM, 64QAM, 256QAM; symbols complex64[n] -> int16 LLRs [n * bits per symbol]"""
        s = np.ascontiguousarray(symbols, np.complex64).view(np.float32)
        n = len(s) // 2
        llr = np.zeros(n * (1, 2, 4, 6, 8)[mod] + 16, np.int16)
        assert self._demod(mod, s, llr, n) == 0
        return llr[:n * (1, 2, 4, 6, 8)[mod]]

    def sequence_apply_s(self, llr, c_init):
        """(de)scrambling of int16 LLRs with the LTE Gold sequence of seed c_init"""
        llr = np.ascontiguousarray(llr, np.int16)
        out = np.zeros_like(llr)
        self._seq(llr, out, len(llr), c_init)
        return out

    def encode_tb(self, tbs, Qm, rv, nof_e_bits, data):
        """-> (ret, e_bits packed MSB-first, (nof_e_bits+7)//8 bytes)"""
        data = np.ascontiguousarray(data, np.uint8)
        e = np.zeros((nof_e_bits + 7) // 8 + 64, np.uint8)
        ret = self._enc_tb(tbs, Qm, rv, nof_e_bits, data, e)
        return ret, e[:(nof_e_bits + 7) // 8]

    # ---- literal sch.c functions (compiled reference only)
    def dlsch_encode(self, tbs, Qm, rv, nof_e_bits, data):
        data = np.ascontiguousarray(data, np.uint8).copy()
        e = np.zeros((nof_e_bits + 7) // 8 + 64, np.uint8)
        ret = self.lib.ref_dlsch_encode(tbs, Qm, rv, nof_e_bits, data, e)
        return ret, e[:(nof_e_bits + 7) // 8]

    def dlsch_rx_new(self):
        return self.lib.ref_dlsch_rx_new()

    def dlsch_rx_new_guru(self, max_cb):
        return self.lib.ref_dlsch_rx_new_guru(max_cb)

    def dlsch_rx_free(self, h):
        self.lib.ref_dlsch_rx_free(h)

    def dlsch_decode(self, h, tbs, Qm, rv, e_bits, max_iterations):
        """srsran_dlsch_decode2 with the production (AUTO) decoder on the persistent soft buffer h"""
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        ncb = self.lib.ref_dlsch_rx_max_cb(h)
        data = np.zeros(ncb * 768 + 8, np.uint8)
        cbc = np.zeros(ncb, np.uint8); tbc = np.zeros(1, np.uint8); avg = C.c_float(0)
        ret = self.lib.ref_dlsch_decode(h, tbs, Qm, rv, len(e_bits), e_bits, max_iterations, data, cbc, tbc, C.byref(avg))
        return dict(ret=ret, data=data, cb_crc=cbc, tb_crc=int(tbc[0]), avg_iterations=avg.value)

    def dlsch_decode_cw(self, h, tbs, Qm, rv, e_bits, max_iterations, tb_idx, nof_layers, nof_tb):
        """srsran_dlsch_decode2(q, cfg, e_bits, data, tb_idx, nof_layers) on a grant with nof_tb codewords; Qm = bits per symbol of the
        modulation (decode_tb sees Qm * Nl, Nl = 2 when nof_layers != nof_tb: sch.c:587-604)"""
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        ncb = self.lib.ref_dlsch_rx_max_cb(h)
        data = np.zeros(ncb * 768 + 8, np.uint8)
        cbc = np.zeros(ncb, np.uint8); tbc = np.zeros(1, np.uint8); avg = C.c_float(0)
        ret = self.lib.ref_dlsch_decode_cw(h, tbs, Qm, rv, len(e_bits), e_bits, max_iterations, data, cbc, tbc, C.byref(avg), tb_idx, nof_layers, nof_tb)
        return dict(ret=ret, data=data, cb_crc=cbc, tb_crc=int(tbc[0]), avg_iterations=avg.value)

    def dlsch_encode_retx_null(self, tbs, Qm, rv, nof_e_bits, data):
        """srsran_dlsch_encode2 with data, then again with data == NULL and redundancy version rv on the same tx soft buffer"""
        data = np.ascontiguousarray(data, np.uint8).copy()
        e = np.zeros((nof_e_bits + 7) // 8 + 64, np.uint8)
        ret = self.lib.ref_dlsch_encode_retx_null(tbs, Qm, rv, nof_e_bits, data, e)
        return ret, e[:(nof_e_bits + 7) // 8]

    def dlsch_encode_cw(self, tbs, Qm, rv, nof_e_bits, data, tb_idx, nof_layers, nof_tb):
        data = np.ascontiguousarray(data, np.uint8).copy()
        e = np.zeros((nof_e_bits + 7) // 8 + 64, np.uint8)
        ret = self.lib.ref_dlsch_encode_cw(tbs, Qm, rv, nof_e_bits, data, e, tb_idx, nof_layers, nof_tb)
        return ret, e[:(nof_e_bits + 7) // 8]

    def dlsch_decode8(self, h, tbs, Qm, rv, e_bits8, max_iterations):
        """srsran_dlsch_decode2 with q->llr_is_8bit set (int8 LLRs) on the persistent soft buffer h"""
        e = np.ascontiguousarray(e_bits8, np.int8)
        ncb = self.lib.ref_dlsch_rx_max_cb(h)
        data = np.zeros(ncb * 768 + 8, np.uint8)
        cbc = np.zeros(ncb, np.uint8); tbc = np.zeros(1, np.uint8)
        ret = self.lib.ref_dlsch_decode8(h, tbs, Qm, rv, len(e), e, max_iterations, data, cbc, tbc)
        return dict(ret=ret, data=data, cb_crc=cbc, tb_crc=int(tbc[0]))

    def ulsch_encode(self, tbs, Qm, rv, nof_symb, L_prb, data, ri_len=0, ri_value=0):
        """srsran_ulsch_encode -> (ret, interleaved hard bits q[L_prb*12*nof_symb*Qm])"""
        nb = L_prb * 12 * nof_symb * Qm
        q = np.zeros(nb // 8 + 64, np.uint8)
        ret = self.lib.ref_ulsch_encode(tbs, Qm, rv, nof_symb, L_prb, ri_len, ri_value, np.ascontiguousarray(data, np.uint8).copy(), q)
        return ret, np.unpackbits(q)[:nb]

    def ulsch_decode(self, h, tbs, Qm, rv, nof_symb, L_prb, q_llr, max_iterations, ri_len=0):
        """srsran_ulsch_decode on the persistent soft buffer h"""
        q_llr = np.ascontiguousarray(q_llr, np.int16)
        ncb = self.lib.ref_dlsch_rx_max_cb(h)
        data = np.zeros(ncb * 768 + 8, np.uint8)
        cbc = np.zeros(ncb, np.uint8); tbc = np.zeros(1, np.uint8); ri = np.zeros(1, np.uint8); avg = C.c_float(0)
        ret = self.lib.ref_ulsch_decode(h, tbs, Qm, rv, nof_symb, L_prb, ri_len, q_llr, max_iterations, data, cbc, tbc, C.byref(avg), ri)
        return dict(ret=ret, data=data, cb_crc=cbc, tb_crc=int(tbc[0]), avg_iterations=avg.value, ri=int(ri[0]))

    def ulsch_deinterleave(self, q_bits, Qm, H_prime_total, N_pusch_symbs, ri_positions=()):
        q_bits = np.ascontiguousarray(q_bits, np.int16)
        g = np.zeros(H_prime_total * Qm + 8, np.int16)
        ri = np.ascontiguousarray(np.array(list(ri_positions) + [0], np.uint32))
        getattr(self.lib, self.p + "ulsch_deinterleave")(q_bits, Qm, H_prime_total, N_pusch_symbs, g, ri, len(ri_positions))
        return g[:H_prime_total * Qm]


def new_tb_state(Cn):
    return dict(buffer_f=np.zeros((Cn, SOFTBUFFER_SIZE), np.int16), sb_data=np.zeros((Cn, SOFTBUFFER_SIZE // 8), np.uint8),
                cb_crc=np.zeros(Cn, np.uint8))


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        _oracle = _Lib(build_oracle(), "orc_")
    return _oracle


PHY_B200_SO = os.path.join(ROOT, "integration", "_build", "libsrsran_phy_b200.so")
_phy = None


def phy_b200():
    """The reference's own sch.c with the SRSRAN_B200 hooks applied, linked against the shim + libsrsran_b200.so
    (integration/Makefile `phy`); same flat entry points as ref(). Needs a GPU. None when it has not been built."""
    global _phy
    if _phy is None and os.path.exists(PHY_B200_SO):
        _phy = _Lib(PHY_B200_SO, "ref_")
    return _phy


def ref():
    """The compiled reference, or None when oracle/_ref has not been (cannot be) built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = _Lib(REF_SO, "ref_")
    return _ref


REF_NATIVE_SO = os.path.join(ORACLE_DIR, "_ref", "libsrsran_ref_native.so")
_ref_native = None


def ref_for_timing():
    """-> (library, build description) for the CPU speed baseline: the reference compiled with its full release flags
    including -march=native (oracle/Makefile) when THIS host has every ISA extension of the build host, else the portable
    AVX2 + FMA build. Parity tests never use this one."""
    global _ref_native
    flags_file = os.path.join(ORACLE_DIR, "_ref", "native_cpu_flags.txt")
    if os.path.exists(REF_NATIVE_SO) and os.path.exists(flags_file):
        need = set(open(flags_file).read().split())
        have = set()
        try:
            for line in open("/proc/cpuinfo"):
                if line.startswith("flags"):
                    have = set(line.split(":", 1)[1].split())
                    break
        except OSError:
            pass
        if need and need <= have:
            if _ref_native is None:
                _ref_native = _Lib(REF_NATIVE_SO, "ref_")
            return _ref_native, "-O3 -Ofast -funroll-loops -march=native (reference release flags)"
    r = ref()
    return r, "-O3 -Ofast -funroll-loops -mavx2 -mfma (host lacks ISA extensions of the build host: no -march=native)"

