#!/usr/bin/env python3
"""Where does the end-to-end step go? srsb200_tdec_batch on the bench workload with different numbers of copy/compute
ranges and with the decode reduced to one half-iteration (copy-bound floor)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsran_4g_b200 as sb  # noqa: E402
from srsran_4g_b200 import synth  # noqa: E402

K, n_cb = 6144, 16384
dev = torch.device("cuda", 0)
eng = sb.Engine(0)
L = sb.lib()
bits, llr = synth.make_llr_batch(K, n_cb, 1.5, 1000, 100, n_distinct=256, device=dev)
h_llr = torch.empty((n_cb, 3 * K + 12), dtype=torch.int16, pin_memory=True)
h_llr.copy_(llr)
h_out = torch.empty((n_cb, K // 8), dtype=torch.uint8, pin_memory=True)
h_noi = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)
h_ok = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)
Ks = np.full(n_cb, K, np.uint32)
kinds = np.full(n_cb, sb.CRC_24B, np.uint8)
loff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(3 * K + 12))
ooff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(K // 8))
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def step(max_iter, early):
    r = L.srsb200_tdec_batch(eng.handle, n_cb, vp(Ks), vp(kinds), C.c_void_p(h_llr.data_ptr()), vp(loff), n_cb * (3 * K + 12), max_iter, 2, early,
                             C.c_void_p(h_out.data_ptr()), vp(ooff), n_cb * (K // 8), C.c_void_p(h_noi.data_ptr()), C.c_void_p(h_ok.data_ptr()))
    assert r == 0


res = {}
for nsub in (1, 2, 4, 8):
    L.srsb200_engine_set_subbatches(eng.handle, nsub)
    for name, (mi, es) in (("full", (8, 1)), ("one_half_iteration", (1, 0))):
        for _ in range(2):
            step(mi, es)
        t0 = time.perf_counter()
        for _ in range(5):
            step(mi, es)
        res[f"nsub{nsub}_{name}"] = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps(res, indent=1))
