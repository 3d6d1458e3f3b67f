"""ctypes binding of libsrsran_b200.so (the C ABI in include/srsran_b200.h)."""
import ctypes as C
import os

import numpy as np

CRC_NONE, CRC_24A, CRC_24B = 0, 1, 2
SOFTBUFFER_SIZE = 18600
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SrsB200Error(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libsrsran_b200.so")


def lib():
    """Load the CUDA library; fails loudly when it has not been built (no silent fallback)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise SrsB200Error("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % p)
        L = C.CDLL(p)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.srsb200_last_error.restype = C.c_char_p
        L.srsb200_engine_create.argtypes = [C.POINTER(vp), i32]
        L.srsb200_engine_destroy.argtypes = [vp]
        L.srsb200_engine_destroy.restype = None
        L.srsb200_engine_launch_count.argtypes = [vp]
        L.srsb200_engine_launch_count.restype = u64
        L.srsb200_engine_stream.argtypes = [vp]
        L.srsb200_engine_stream.restype = vp
        L.srsb200_engine_sync.argtypes = [vp]
        L.srsb200_engine_flush.argtypes = [vp]
        L.srsb200_host_alloc.argtypes = [C.c_size_t]
        L.srsb200_host_alloc.restype = vp
        L.srsb200_host_free.argtypes = [vp]
        L.srsb200_host_register.argtypes = [vp, C.c_size_t]
        L.srsb200_host_unregister.argtypes = [vp]
        L.srsb200_engine_profile.argtypes = [vp, i32]
        L.srsb200_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i32]
        L.srsb200_multi_destroy.argtypes = [vp]; L.srsb200_multi_destroy.restype = None
        L.srsb200_multi_nof_devices.argtypes = [vp]
        L.srsb200_multi_engine.argtypes = [vp, i32]; L.srsb200_multi_engine.restype = vp
        L.srsb200_multi_device_of.argtypes = [vp, u64]
        L.srsb200_multi_decode_tb_batch.argtypes = [vp, vp, u32, vp, u32]
        L.srsb200_multi_tdec_batch.argtypes = [vp, u32, vp, vp, vp, vp, u64, u32, u32, i32, vp, vp, u64, vp, vp]
        L.srsb200_demod_soft_demodulate_s.argtypes = [vp, u32, vp, vp, u32]
        L.srsb200_tdec8_windows.argtypes = [u32]; L.srsb200_tdec8_windows.restype = u32
        L.srsb200_tdec_batch8.argtypes = [vp, u32, vp, vp, vp, vp, u64, u32, u32, i32, vp, vp, u64, vp, vp]
        L.srsb200_rm_turbo_rx_lut8.argtypes = [vp, vp, vp, u32, u32, u32]
        L.srsb200_engine_set_subbatches.argtypes = [vp, i32]
        L.srsb200_engine_inject_alloc_failure.argtypes = [vp, i32]
        L.srsb200_engine_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
        L.srsb200_engine_profile_read_kinds.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64), C.c_uint32]
        L.srsb200_cbsize.argtypes = [u32]
        L.srsb200_cbindex.argtypes = [u32]
        L.srsb200_cbsegm.argtypes = [u32, C.POINTER(u32)]
        L.srsb200_tdec_autoimp_get_subblocks.argtypes = [u32]
        L.srsb200_tdec_autoimp_get_subblocks.restype = u32
        L.srsb200_tdec_batch.argtypes = [vp, u32, vp, vp, vp, vp, u64, u32, u32, i32, vp, vp, u64, vp, vp]
        L.srsb200_tdec_plan_uniform.argtypes = [vp, u32, u32, i32, C.POINTER(vp)]
        L.srsb200_plan_destroy.argtypes = [vp]
        L.srsb200_plan_regroup_points.argtypes = [vp, vp, vp, C.c_uint32]
        L.srsb200_plan_destroy.restype = None
        L.srsb200_tdec_run_plan_dev.argtypes = [vp, vp, vp, u32, u32, i32, vp, vp, vp]
        L.srsb200_tdec_init.argtypes = [C.POINTER(vp), vp, u32]
        L.srsb200_tdec_free.argtypes = [vp]
        L.srsb200_tdec_free.restype = None
        L.srsb200_tdec_new_cb.argtypes = [vp, u32]
        L.srsb200_tdec_get_nof_iterations.argtypes = [vp]
        L.srsb200_tdec_iteration.argtypes = [vp, vp, vp]
        L.srsb200_tdec_run_all.argtypes = [vp, vp, vp, u32, u32]
        L.srsb200_rm_turbo_gentables.argtypes = [vp]
        L.srsb200_rm_turbo_rx_lut.argtypes = [vp, vp, vp, u32, u32, u32]
        L.srsb200_rm_table.argtypes = [u32, u32, vp]
        L.srsb200_decode_tb_batch.argtypes = [vp, vp, u32, u32]
        L.srsb200_softbuffer_set_resident.argtypes = [vp, i32]
        for fn in (L.srsb200_softbuffer_reset, L.srsb200_softbuffer_sync_to_host, L.srsb200_softbuffer_release):
            fn.argtypes = [vp, C.POINTER(C.c_void_p), u32]
        L.srsb200_decode_tb.argtypes = [vp, vp, u32]
        L.srsb200_encode_tb_batch.argtypes = [vp, vp, u32]
        L.srsb200_ulsch_deinterleave.argtypes = [vp, vp, u32, u32, u32, vp, vp, u32]
        L.srsb200_encode_tb.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _check(ret, what):
    if ret != 0:
        raise SrsB200Error("%s failed (%d): %s" % (what, ret, lib().srsb200_last_error().decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- host-side metadata (no device needed)
def cbsize(idx):
    return lib().srsb200_cbsize(idx)


def tdec8_windows(K):
    return int(lib().srsb200_tdec8_windows(K))


def cbindex(K):
    return lib().srsb200_cbindex(K)


def cbsegm(tbs):
    out = (C.c_uint32 * 8)()
    ret = lib().srsb200_cbsegm(tbs, out)
    return ret, dict(zip(["F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2"], [int(v) for v in out]))


def rm_table(cb_idx, rv):
    K = cbsize(cb_idx)
    t = np.zeros(3 * K + 12, np.uint16)
    _check(lib().srsb200_rm_table(cb_idx, rv, _ptr(t)), "srsb200_rm_table")
    return t


class _TbStruct(C.Structure):
    _fields_ = [("tbs", C.c_uint32), ("Qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32), ("e_bits", C.c_void_p),
                ("buffer_f", C.POINTER(C.c_void_p)), ("sb_data", C.POINTER(C.c_void_p)), ("cb_crc", C.c_void_p), ("tb_crc", C.c_void_p),
                ("max_cb", C.c_uint32), ("data", C.c_void_p), ("cb_noi", C.c_void_p), ("avg_iterations", C.c_float), ("ret", C.c_int),
                ("q_bits", C.c_void_p), ("H_prime_total", C.c_uint32), ("N_pusch_symbs", C.c_uint32), ("ri_positions", C.c_void_p),
                ("nof_ri_bits", C.c_uint32), ("e_offset", C.c_uint32), ("g_bits", C.c_void_p), ("nof_g_out", C.c_uint32),
                ("descramble", C.c_uint32), ("c_init", C.c_uint32), ("max_iterations", C.c_uint32), ("llr_is_8bit", C.c_uint32),
                ("symbols", C.c_void_p), ("nof_symbols", C.c_uint32), ("mod", C.c_uint32), ("q_gather_pos", C.c_void_p),
                ("q_gather_out", C.c_void_p), ("nof_q_gather", C.c_uint32)]


class _TbTxStruct(C.Structure):
    _fields_ = [("tbs", C.c_uint32), ("Qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32), ("max_cb", C.c_uint32),
                ("data", C.c_void_p), ("e_bits", C.c_void_p), ("ret", C.c_int32)]


class TransportBlock:
    """One decode_tb request + its HARQ soft buffer (srsran_softbuffer_rx_t: buffer_f / data / cb_crc / tb_crc)."""

    def __init__(self, tbs, max_cb=None):
        _, seg = cbsegm(tbs)
        self.tbs, self.seg = tbs, seg
        self.max_cb = max_cb if max_cb is not None else max(seg["C"], 1)
        self.buffer_f = np.zeros((self.max_cb, SOFTBUFFER_SIZE), np.int16)
        self.sb_data = np.zeros((self.max_cb, SOFTBUFFER_SIZE // 8), np.uint8)
        self.cb_crc = np.zeros(self.max_cb, np.uint8)
        self.tb_crc = np.zeros(1, np.uint8)
        self.data = np.zeros(self.max_cb * 768 + 8, np.uint8)
        self.cb_noi = np.zeros(self.max_cb, np.uint32)
        self._bf = (C.c_void_p * self.max_cb)(*[self.buffer_f[i].ctypes.data for i in range(self.max_cb)])
        self._sd = (C.c_void_p * self.max_cb)(*[self.sb_data[i].ctypes.data for i in range(self.max_cb)])
        self.ret, self.avg_iterations = None, 0.0
        self._e = None

    @property
    def buffer_b(self):
        """the soft buffers as the 8-bit mode sees them: (int8_t*)softbuffer->buffer_f[r] (sch.c:410)"""
        return self.buffer_f.view(np.int8)

    def fill(self, s, Qm, rv, e_bits, nof_e_bits=None, llr8=False):
        self._e = np.ascontiguousarray(e_bits, np.int8 if llr8 else np.int16)
        s.llr_is_8bit = 1 if llr8 else 0
        s.tbs, s.Qm, s.rv = self.tbs, Qm, rv
        s.nof_e_bits = len(self._e) if nof_e_bits is None else nof_e_bits
        s.e_bits = self._e.ctypes.data
        s.buffer_f, s.sb_data = self._bf, self._sd
        s.cb_crc, s.tb_crc = self.cb_crc.ctypes.data, self.tb_crc.ctypes.data
        s.max_cb = self.max_cb
        s.data, s.cb_noi = self.data.ctypes.data, self.cb_noi.ctypes.data

    def fill_symbols(self, s, Qm, rv, symbols, mod, nof_e_bits, c_init=None, ul=None, q_gather=()):
        """symbol source: the soft demodulator (and the descrambling) run on the device. ul = dict(H_prime_total, N_pusch_symbs,
        ri_positions, e_offset) for the uplink chain (de-interleaver in between); q_gather: positions of q returned to the host"""
        self.fill(s, Qm, rv, np.zeros(0, np.int16), nof_e_bits)
        s.e_bits = None
        self._sym = np.ascontiguousarray(symbols, np.complex64)
        s.symbols, s.nof_symbols, s.mod = self._sym.ctypes.data, len(self._sym), mod
        if c_init is not None:
            s.descramble, s.c_init = 1, c_init
        if ul is not None:
            self._ri = np.ascontiguousarray(np.array(list(ul.get("ri_positions", ())), np.uint32))
            s.H_prime_total, s.N_pusch_symbs = ul["H_prime_total"], ul["N_pusch_symbs"]
            s.ri_positions = self._ri.ctypes.data if len(self._ri) else None
            s.nof_ri_bits, s.e_offset = len(self._ri), ul.get("e_offset", 0)
            self._gp = np.ascontiguousarray(np.array(list(q_gather), np.uint32))
            self.q_gather_out = np.zeros(max(len(self._gp), 1), np.int16)
            if len(self._gp):
                s.q_gather_pos, s.q_gather_out, s.nof_q_gather = self._gp.ctypes.data, self.q_gather_out.ctypes.data, len(self._gp)

    def fill_ul(self, s, Qm, rv, q_bits, H_prime_total, N_pusch_symbs, nof_e_bits, ri_positions=(), e_offset=0, nof_g_out=0):
        """UL-SCH source: the e-bits are produced on the device by the channel de-interleaver from q_bits"""
        self.fill(s, Qm, rv, np.zeros(0, np.int16), nof_e_bits)
        s.e_bits = None
        self._q = np.ascontiguousarray(q_bits, np.int16)
        self._ri = np.ascontiguousarray(np.array(list(ri_positions), np.uint32))
        self.g_bits = np.zeros(max(nof_g_out, 1), np.int16)
        s.q_bits, s.H_prime_total, s.N_pusch_symbs = self._q.ctypes.data, H_prime_total, N_pusch_symbs
        s.ri_positions = self._ri.ctypes.data if len(self._ri) else None
        s.nof_ri_bits, s.e_offset = len(self._ri), e_offset
        s.g_bits, s.nof_g_out = self.g_bits.ctypes.data, nof_g_out


class Engine:
    """One engine per GPU (owns a stream, device tables and workspaces)."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        self._L = lib()
        _check(self._L.srsb200_engine_create(C.byref(self._h), device), "srsb200_engine_create")

    def close(self):
        if self._h:
            self._L.srsb200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self):
        return int(self._L.srsb200_engine_launch_count(self._h))

    @property
    def stream(self):
        return self._L.srsb200_engine_stream(self._h)

    def sync(self):
        _check(self._L.srsb200_engine_sync(self._h), "srsb200_engine_sync")

    def flush(self):
        """order the engine stream after every device-resident submission still in flight (does not block the host)"""
        _check(self._L.srsb200_engine_flush(self._h), "srsb200_engine_flush")

    def set_subbatches(self, n):
        _check(self._L.srsb200_engine_set_subbatches(self._h, n), "srsb200_engine_set_subbatches")

    def inject_alloc_failure(self, nth):
        _check(self._L.srsb200_engine_inject_alloc_failure(self._h, nth), "srsb200_engine_inject_alloc_failure")

    def profile(self, enable):
        _check(self._L.srsb200_engine_profile(self._h, int(enable)), "srsb200_engine_profile")

    def profile_read(self):
        """-> {kernel kind: (summed ms, launches)} since the last read; kinds: extract, decode, emit, rm, tbcrc"""
        ms = (C.c_double * 11)()
        cnt = (C.c_uint64 * 11)()
        _check(self._L.srsb200_engine_profile_read_kinds(self._h, ms, cnt, 11), "srsb200_engine_profile_read_kinds")
        names = ["extract", "decode", "emit", "rm", "tbcrc", "scan", "job", "tbenc", "deint", "demod", "regroup"]
        return {n: (ms[i], int(cnt[i])) for i, n in enumerate(names)}

    # ---- batched decode, host buffers
    def tdec_batch(self, K, llr, max_iter, early_stop=True, min_iter=2, crc_kind=CRC_24B):
        """K: int or per-block sizes; llr: [n, 3K+12] int16 (uniform K) or list of int16 arrays.
        -> (out list/array of K/8-byte rows, noi[n], crc_ok[n])"""
        if np.isscalar(K):
            llr = np.ascontiguousarray(llr, np.int16)
            n = llr.shape[0]
            Ks = np.full(n, K, np.uint32)
            flat = llr.reshape(-1)
            loff = np.arange(n, dtype=np.uint64) * np.uint64(3 * K + 12)
        else:
            Ks = np.ascontiguousarray(K, np.uint32)
            n = len(Ks)
            parts = [np.ascontiguousarray(a, np.int16).reshape(-1) for a in llr]
            lens = np.array([len(p) for p in parts], np.uint64)
            loff = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
            flat = np.concatenate(parts) if n else np.zeros(0, np.int16)
        kinds = np.full(n, crc_kind, np.uint8) if np.isscalar(crc_kind) else np.ascontiguousarray(crc_kind, np.uint8)
        obytes = (Ks // 8).astype(np.uint64)
        ooff = np.concatenate([[0], np.cumsum(obytes)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        out = np.zeros(int(obytes.sum()), np.uint8)
        noi = np.zeros(n, np.uint8)
        ok = np.zeros(n, np.uint8)
        _check(self._L.srsb200_tdec_batch(self._h, n, _ptr(Ks), _ptr(kinds), _ptr(flat), _ptr(loff), len(flat), max_iter, min_iter,
                                          int(early_stop), _ptr(out), _ptr(ooff), len(out), _ptr(noi), _ptr(ok)), "srsb200_tdec_batch")
        if np.isscalar(K):
            return out.reshape(n, K // 8), noi, ok
        return [out[int(o):int(o) + int(b)] for o, b in zip(ooff, obytes)], noi, ok

    def tdec_batch8(self, K, llr8, max_iter, early_stop=True, min_iter=2, crc_kind=CRC_24B):
        """8-bit LLR mode, uniform K: llr8 [n, 3K+12] int8 -> (out [n, K/8], noi, crc_ok)"""
        llr8 = np.ascontiguousarray(llr8, np.int8)
        n = llr8.shape[0]
        Ks = np.full(n, K, np.uint32); kinds = np.full(n, crc_kind, np.uint8)
        loff = np.arange(n, dtype=np.uint64) * np.uint64(3 * K + 12)
        ooff = np.arange(n, dtype=np.uint64) * np.uint64(K // 8)
        out = np.zeros((n, K // 8), np.uint8); noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
        _check(self._L.srsb200_tdec_batch8(self._h, n, _ptr(Ks), _ptr(kinds), _ptr(llr8), _ptr(loff), llr8.size, max_iter, min_iter, int(early_stop),
                                           _ptr(out), _ptr(ooff), out.size, _ptr(noi), _ptr(ok)), "srsb200_tdec_batch8")
        return out, noi, ok

    def tdec_batch8_mixed(self, Ks, llrs8, max_iter, early_stop=True, min_iter=2, crc_kind=CRC_24B):
        """8-bit LLR mode, one block size per entry: Ks uint32[n], llrs8 list of int8 arrays (3K+12 each) -> (list of K/8-byte
        rows, noi, crc_ok)"""
        Ks = np.ascontiguousarray(Ks, np.uint32)
        n = len(Ks)
        parts = [np.ascontiguousarray(a, np.int8).reshape(-1) for a in llrs8]
        flat = np.concatenate(parts) if n else np.zeros(0, np.int8)
        loff = np.concatenate([[0], np.cumsum([len(p) for p in parts])[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        ooff = np.concatenate([[0], np.cumsum(Ks.astype(np.uint64) // 8)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        kinds = np.full(n, crc_kind, np.uint8) if np.isscalar(crc_kind) else np.ascontiguousarray(crc_kind, np.uint8)
        out = np.zeros(int((Ks.astype(np.uint64) // 8).sum()), np.uint8); noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
        _check(self._L.srsb200_tdec_batch8(self._h, n, _ptr(Ks), _ptr(kinds), _ptr(flat), _ptr(loff), flat.size, max_iter, min_iter, int(early_stop),
                                           _ptr(out), _ptr(ooff), out.size, _ptr(noi), _ptr(ok)), "srsb200_tdec_batch8")
        return [out[int(ooff[i]):int(ooff[i]) + int(Ks[i]) // 8] for i in range(n)], noi, ok

    def demod_soft_demodulate_s(self, mod, symbols):
        """mod 0..4 = BPSK, QPSK, 16QAM, 64QAM, 256QAM; symbols complex64[n] -> (ret, int16 LLRs)"""
        s = np.ascontiguousarray(symbols, np.complex64)
        llr = np.zeros(len(s) * (1, 2, 4, 6, 8)[mod if mod < 5 else 0], np.int16)
        ret = self._L.srsb200_demod_soft_demodulate_s(self._h, mod, _ptr(s), _ptr(llr), len(s))
        return ret, llr

    def rm_turbo_rx_lut8(self, e8, buf8, cb_idx, rv):
        e8 = np.ascontiguousarray(e8, np.int8)
        return self._L.srsb200_rm_turbo_rx_lut8(self._h, _ptr(e8), _ptr(buf8), len(e8), cb_idx, rv)

    # ---- device-resident plan API (raw device pointers, e.g. torch tensor .data_ptr())
    def plan_uniform(self, n, K, crc_kind=CRC_24B):
        p = C.c_void_p()
        _check(self._L.srsb200_tdec_plan_uniform(self._h, n, K, crc_kind, C.byref(p)), "srsb200_tdec_plan_uniform")
        return p

    def plan_destroy(self, p):
        self._L.srsb200_plan_destroy(p)

    def plan_regroup_points(self, p):
        """per range of the plan's last device-resident decode: the half-iteration count after which its unfinished blocks were
        regrouped (0: not); [] for plans that never regroup"""
        pts = (C.c_uint32 * 16)()
        r = self._L.srsb200_plan_regroup_points(self._h, p, pts, 16)
        if r < 0:
            _check(r, "srsb200_plan_regroup_points")
        return list(pts)[:r]

    def run_plan_dev(self, plan, d_llr, max_iter, min_iter, early_stop, d_out, d_noi, d_ok):
        _check(self._L.srsb200_tdec_run_plan_dev(self._h, plan, d_llr, max_iter, min_iter, int(early_stop), d_out, d_noi, d_ok),
               "srsb200_tdec_run_plan_dev")

    # ---- rate de-matching (srsran_rm_turbo_rx_lut_ with natural layout)
    def rm_turbo_rx_lut(self, e, buf, cb_idx, rv):
        e = np.ascontiguousarray(e, np.int16)
        assert buf.dtype == np.int16 and buf.flags.c_contiguous
        return self._L.srsb200_rm_turbo_rx_lut(self._h, _ptr(e), _ptr(buf), len(e), cb_idx, rv)

    # ---- HARQ soft buffers: host-coherent (default) or device-resident mirror
    def softbuffer_set_resident(self, on):
        _check(self._L.srsb200_softbuffer_set_resident(self._h, int(on)), "srsb200_softbuffer_set_resident")

    def softbuffer_reset(self, tb):
        """srsran_softbuffer_rx_reset for a TransportBlock in resident mode (also clears its host-side flags)"""
        tb.cb_crc[:] = 0  # (the host buffer_f is not authoritative in resident mode and is left alone)
        tb.tb_crc[:] = 0
        _check(self._L.srsb200_softbuffer_reset(self._h, tb._bf, tb.max_cb), "srsb200_softbuffer_reset")

    def softbuffer_sync_to_host(self, tb):
        _check(self._L.srsb200_softbuffer_sync_to_host(self._h, tb._bf, tb.max_cb), "srsb200_softbuffer_sync_to_host")

    def softbuffer_release(self, tb):
        _check(self._L.srsb200_softbuffer_release(self._h, tb._bf, tb.max_cb), "srsb200_softbuffer_release")

    # ---- transport blocks
    def decode_tb_batch(self, reqs, max_iterations, limits=None, llr8=False):
        """reqs: list of (TransportBlock, Qm, rv, e_bits); limits: optional per-TB half-iteration limits (0 = max_iterations);
        llr8: q->llr_is_8bit - int8 e-bits and soft buffers, the reference's windowed saturating decoders"""
        arr = (_TbStruct * len(reqs))()
        for i, (s, (tb, Qm, rv, e)) in enumerate(zip(arr, reqs)):
            tb.fill(s, Qm, rv, e, llr8=llr8)
            if limits is not None:
                s.max_iterations = int(limits[i])
        ret = self._L.srsb200_decode_tb_batch(self._h, arr, len(reqs), max_iterations)
        for s, (tb, _, _, _) in zip(arr, reqs):
            tb.ret, tb.avg_iterations = s.ret, s.avg_iterations
        return ret

    def encode_tb_batch(self, reqs, max_cb=256):
        """reqs: list of (tbs, Qm, rv, nof_e_bits, data bytes) -> (ret, [(tb_ret, e_bits packed uint8)])"""
        arr = (_TbTxStruct * len(reqs))()
        keep = []
        for s, (tbs, Qm, rv, G, data) in zip(arr, reqs):
            d = None if data is None else np.ascontiguousarray(data, np.uint8)
            e = np.zeros((G + 7) // 8, np.uint8)
            keep.append((d, e))
            s.tbs, s.Qm, s.rv, s.nof_e_bits, s.max_cb = tbs, Qm, rv, G, max_cb
            s.data = None if d is None else d.ctypes.data
            s.e_bits = e.ctypes.data
        ret = self._L.srsb200_encode_tb_batch(self._h, arr, len(reqs))
        return ret, [(int(s.ret), k[1]) for s, k in zip(arr, keep)]

    def encode_tb(self, tbs, Qm, rv, nof_e_bits, data, max_cb=256):
        s = _TbTxStruct()
        d = None if data is None else np.ascontiguousarray(data, np.uint8)
        e = np.zeros((nof_e_bits + 7) // 8, np.uint8)
        s.tbs, s.Qm, s.rv, s.nof_e_bits, s.max_cb = tbs, Qm, rv, nof_e_bits, max_cb
        s.data = None if d is None else d.ctypes.data
        s.e_bits = e.ctypes.data
        ret = self._L.srsb200_encode_tb(self._h, C.byref(s))
        return ret, e

    def ulsch_deinterleave(self, q_bits, Qm, H_prime_total, N_pusch_symbs, ri_positions=()):
        q = np.ascontiguousarray(q_bits, np.int16)
        g = np.zeros(H_prime_total * Qm, np.int16)
        ri = np.ascontiguousarray(np.array(list(ri_positions), np.uint32))
        ret = self._L.srsb200_ulsch_deinterleave(self._h, _ptr(q), Qm, H_prime_total, N_pusch_symbs, _ptr(g), _ptr(ri) if len(ri) else None, len(ri))
        return ret, g

    def decode_tb_symbols(self, tb, Qm, rv, symbols, mod, nof_e_bits, max_iterations, c_init=None, ul=None, q_gather=()):
        s = _TbStruct()
        tb.fill_symbols(s, Qm, rv, symbols, mod, nof_e_bits, c_init, ul, q_gather)
        ret = self._L.srsb200_decode_tb(self._h, C.byref(s), max_iterations)
        tb.ret, tb.avg_iterations = s.ret, s.avg_iterations
        return ret

    def ulsch_decode_batch(self, reqs, max_iterations):
        """reqs: list of (TransportBlock, Qm, rv, q_bits, H_prime_total, N_pusch_symbs, nof_e_bits, ri_positions, e_offset, nof_g_out)"""
        arr = (_TbStruct * len(reqs))()
        for s, r in zip(arr, reqs):
            r[0].fill_ul(s, *r[1:])
        ret = self._L.srsb200_decode_tb_batch(self._h, arr, len(reqs), max_iterations)
        for s, r in zip(arr, reqs):
            r[0].ret, r[0].avg_iterations = s.ret, s.avg_iterations
        return ret

    def decode_tb(self, tb, Qm, rv, e_bits, max_iterations, nof_e_bits=None, c_init=None, llr8=False):
        """c_init: the e_bits are still scrambled; descramble them on the device with the Gold sequence of that seed;
        llr8: q->llr_is_8bit (int8 e-bits and soft buffers)"""
        s = _TbStruct()
        tb.fill(s, Qm, rv, e_bits, nof_e_bits, llr8=llr8)
        if c_init is not None:
            s.descramble, s.c_init = 1, c_init
        ret = self._L.srsb200_decode_tb(self._h, C.byref(s), max_iterations)
        tb.ret, tb.avg_iterations = s.ret, s.avg_iterations
        return ret


class Tdec:
    """Mirror of srsran_tdec_t + srsran_tdec_* (lib/include/srsran/phy/fec/turbo/turbodecoder.h:63-116)."""

    def __init__(self, engine, max_long_cb=6144):
        self._L = lib()
        self._h = C.c_void_p()
        self.engine = engine
        _check(self._L.srsb200_tdec_init(C.byref(self._h), engine.handle, max_long_cb), "srsb200_tdec_init")

    def free(self):
        if self._h:
            self._L.srsb200_tdec_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def new_cb(self, long_cb):
        return self._L.srsb200_tdec_new_cb(self._h, long_cb)

    def get_nof_iterations(self):
        return self._L.srsb200_tdec_get_nof_iterations(self._h)

    def iteration(self, llr, K):
        llr = np.ascontiguousarray(llr, np.int16)
        out = np.zeros(K // 8, np.uint8)
        _check(self._L.srsb200_tdec_iteration(self._h, _ptr(llr), _ptr(out)), "srsb200_tdec_iteration")
        return out

    def run_all(self, llr, nof_iterations, K):
        llr = np.ascontiguousarray(llr, np.int16)
        out = np.zeros(K // 8, np.uint8)
        ret = self._L.srsb200_tdec_run_all(self._h, _ptr(llr), _ptr(out), nof_iterations, K)
        return ret, out


class Multi:
    """srsb200_multi_*: one process, several GPUs; transport blocks are placed by owner key modulo the device count"""

    def __init__(self, devices=None):
        self._L = lib()
        self._h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            _check(self._L.srsb200_multi_create(C.byref(self._h), arr, len(devices)), "srsb200_multi_create")
        else:
            _check(self._L.srsb200_multi_create(C.byref(self._h), None, 0), "srsb200_multi_create")

    def close(self):
        if self._h:
            self._L.srsb200_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nof_devices(self):
        return self._L.srsb200_multi_nof_devices(self._h)

    def launch_counts(self):
        return [int(self._L.srsb200_engine_launch_count(self._L.srsb200_multi_engine(self._h, i))) for i in range(self.nof_devices)]

    def device_of(self, owner):
        return self._L.srsb200_multi_device_of(self._h, owner)

    def decode_tb_batch(self, reqs, max_iterations, owners=None):
        """reqs: list of (TransportBlock, Qm, rv, e_bits); owners: optional per-TB owner keys (cell ids)"""
        arr = (_TbStruct * len(reqs))()
        for s, (tb, Qm, rv, e) in zip(arr, reqs):
            tb.fill(s, Qm, rv, e)
        own = np.ascontiguousarray(owners, np.uint64) if owners is not None else None
        ret = self._L.srsb200_multi_decode_tb_batch(self._h, arr, len(reqs), _ptr(own) if own is not None else None, max_iterations)
        for s, (tb, _, _, _) in zip(arr, reqs):
            tb.ret, tb.avg_iterations = s.ret, s.avg_iterations
        return ret

    def tdec_batch(self, K, llr, max_iter, early_stop=True, min_iter=2, crc_kind=CRC_24B):
        """uniform K: llr [n, 3K+12] int16 -> (out [n, K/8], noi, ok), split over the devices"""
        llr = np.ascontiguousarray(llr, np.int16)
        n = llr.shape[0]
        Ks = np.full(n, K, np.uint32); kinds = np.full(n, crc_kind, np.uint8)
        loff = np.arange(n, dtype=np.uint64) * np.uint64(3 * K + 12)
        ooff = np.arange(n, dtype=np.uint64) * np.uint64(K // 8)
        out = np.zeros((n, K // 8), np.uint8); noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
        _check(self._L.srsb200_multi_tdec_batch(self._h, n, _ptr(Ks), _ptr(kinds), _ptr(llr), _ptr(loff), llr.size, max_iter, min_iter, int(early_stop),
                                                _ptr(out), _ptr(ooff), out.size, _ptr(noi), _ptr(ok)), "srsb200_multi_tdec_batch")
        return out, noi, ok
