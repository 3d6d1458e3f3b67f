#!/usr/bin/env python3
"""SRSB200_TRACE of the 64-cell subframe submission (config 5), device-resident soft buffers."""
import os, sys, time
os.environ["SRSB200_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import srsran_4g_b200 as sb
import vecgen
tbs, G, Qm, cells = 75376, 86400, 6, 64
el = [vecgen.make_tb(tbs, G, Qm, 0, 6.0, 500 + c, scale=700)[1] for c in range(4)]
eng = sb.Engine(0)
for resident in (True, False):
    eng.softbuffer_set_resident(resident)
    tbl = [sb.TransportBlock(tbs) for _ in range(cells)]
    for rep in range(4):
        t0 = time.perf_counter()
        rq = []
        for c in range(cells):
            tb = tbl[c]
            if resident:
                eng.softbuffer_reset(tb)
            else:
                tb.buffer_f[:] = 0; tb.cb_crc[:] = 0
            rq.append((tb, Qm, 0, el[c % 4]))
        t1 = time.perf_counter()
        eng.decode_tb_batch(rq, 8)
        t2 = time.perf_counter()
        print("resident=%s rep %d: reset %.3f ms, decode_tb_batch %.3f ms" % (resident, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
