#!/usr/bin/env python3
"""SRSB200_TRACE timeline of one end-to-end submission of the bench workload + host wall clock."""
import ctypes as C, os, sys, time
import numpy as np, torch
os.environ["SRSB200_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsran_4g_b200 as sb
from srsran_4g_b200 import synth
K, n_cb = 6144, 16384
dev = torch.device("cuda", 0)
eng = sb.Engine(0); L = sb.lib()
bits, llr = synth.make_llr_batch(K, n_cb, 1.5, 1000, 100, n_distinct=256, device=dev)
h_llr = torch.empty((n_cb, 3 * K + 12), dtype=torch.int16, pin_memory=True); h_llr.copy_(llr)
h_out = torch.empty((n_cb, K // 8), dtype=torch.uint8, pin_memory=True)
h_noi = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True); h_ok = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)
Ks = np.full(n_cb, K, np.uint32); kinds = np.full(n_cb, sb.CRC_24B, np.uint8)
loff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(3 * K + 12)); ooff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(K // 8))
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for i in range(4):
    t0 = time.perf_counter()
    r = L.srsb200_tdec_batch(eng.handle, n_cb, vp(Ks), vp(kinds), C.c_void_p(h_llr.data_ptr()), vp(loff), n_cb * (3 * K + 12), 8, 2, 1,
                             C.c_void_p(h_out.data_ptr()), vp(ooff), n_cb * (K // 8), C.c_void_p(h_noi.data_ptr()), C.c_void_p(h_ok.data_ptr()))
    print("call %d: %.3f ms wall" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
