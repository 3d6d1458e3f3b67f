#!/usr/bin/env python3
"""Where does the mixed-size submission (config 4: 188 sizes x n blocks) spend its time?"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
import srsran_4g_b200 as sb
from srsran_4g_b200 import synth
o = ol.oracle()
eng = sb.Engine(0); L = sb.lib()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Ks, llrs = [], []
for idx in range(188):
    k = o.cbsize(idx)
    _, l = synth.make_llr_batch(k, per, 2.0 if k < 512 else 1.5, 100 + idx, n_distinct=min(per, 8))
    l = np.asarray(l.cpu()) if hasattr(l, "cpu") else np.asarray(l)
    for i in range(per):
        Ks.append(k); llrs.append(l[i])
Ks = np.array(Ks, np.uint32)
n = len(Ks)
flat = np.concatenate(llrs).astype(np.int16)
loff = np.concatenate([[0], np.cumsum(3 * Ks.astype(np.uint64) + 12)[:-1]]).astype(np.uint64)
ooff = np.concatenate([[0], np.cumsum(Ks.astype(np.uint64) // 8)[:-1]]).astype(np.uint64)
out = np.zeros(int((Ks // 8).sum()), np.uint8)
noi = np.zeros(n, np.uint8); ok = np.zeros(n, np.uint8)
kinds = np.full(n, sb.CRC_24B, np.uint8)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
def call():
    r = L.srsb200_tdec_batch(eng.handle, n, vp(Ks), vp(kinds), vp(flat), vp(loff), len(flat), 8, 2, 1, vp(out), vp(ooff), len(out), vp(noi), vp(ok))
    assert r == 0
for _ in range(3):
    call()
t0 = time.perf_counter()
for _ in range(5):
    call()
dt = (time.perf_counter() - t0) / 5
eng.profile(True); eng.profile_read()
call()
prof = eng.profile_read(); eng.profile(False)
print("blocks %d, info bits %d, C-ABI call %.3f ms (%.1f Mbit/s), crc ok %.3f, mean noi %.2f" % (n, Ks.sum(), dt * 1e3, Ks.sum() / dt / 1e6, ok.mean(), noi.mean()))
print({k: (round(v[0], 3), v[1]) for k, v in prof.items() if v[1]})
