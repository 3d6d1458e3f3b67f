#!/usr/bin/env python3
"""Achieved bandwidth of the kernels around the decoder (rate de-matching, transport-block CRC, UL-SCH de-interleaver) on a
large transport-block submission, by CUDA events on the launching stream (engine profiling). Writes one JSON line."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import srsran_4g_b200 as sb
import vecgen

n_tb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tbs, G, Qm = 75376, 86400, 6
C, K = 13, 5824
L = 3 * K + 12
els = [vecgen.make_tb(tbs, G, Qm, 0, 6.0, 500 + c, scale=700)[1] for c in range(4)]
eng = sb.Engine(0)
eng.softbuffer_set_resident(True)
tbl = [sb.TransportBlock(tbs) for _ in range(n_tb)]
def run(ul):
    reqs = []
    for c, tb in enumerate(tbl):
        eng.softbuffer_reset(tb)
        if ul:
            reqs.append((tb, Qm, 0, els[c % 4], G // Qm, 12, G, (), 0, 0))   # q_bits = e_bits: only the de-interleaver's cost matters here
        else:
            reqs.append((tb, Qm, 0, els[c % 4]))
    return eng.ulsch_decode_batch(reqs, 2) if ul else eng.decode_tb_batch(reqs, 2)
out = {"workload": "%d TB x TBS %d (13 CB, K=5824), G=%d, device-resident soft buffers" % (n_tb, tbs, G), "peak_GBps": 6542.1}
for ul in (False, True):
    run(ul); run(ul)
    eng.profile(True); eng.profile_read()
    run(ul)
    p = eng.profile_read(); eng.profile(False)
    if not ul:
        E = G // C
        rm_bytes = n_tb * C * (2 * E + 2 * 2 * L)          # read e-bits, read-modify-write the soft buffer
        crc_bytes = n_tb * (tbs + 24) // 8
        out["rm_rx_kernel"] = {"ms": p["rm"][0], "algorithmic_MB": rm_bytes / 1e6, "GBps": rm_bytes / p["rm"][0] / 1e6, "frac_of_hbm_peak": rm_bytes / p["rm"][0] / 1e6 / 6542.1}
        out["crc_bytes_kernel"] = {"ms": p["tbcrc"][0], "algorithmic_MB": crc_bytes / 1e6, "GBps": crc_bytes / p["tbcrc"][0] / 1e6}
        out["rm_ms_launches"] = p["rm"]
    else:
        de_bytes = n_tb * G * 2 * 2
        de_ms = p["deint"][0]
        out["ulsch_deint_kernel"] = {"ms": de_ms, "algorithmic_MB": de_bytes / 1e6, "GBps": de_bytes / max(de_ms, 1e-6) / 1e6, "frac_of_hbm_peak": de_bytes / max(de_ms, 1e-6) / 1e6 / 6542.1}
# descrambling fused into the rate de-matcher: same submission with descramble / c_init set on every TB
import ctypes as C
from srsran_4g_b200.binding import _TbStruct
def run_scr():
    arr = (_TbStruct * n_tb)()
    for c, (s, tb) in enumerate(zip(arr, tbl)):
        eng.softbuffer_reset(tb)
        tb.fill(s, Qm, 0, els[c % 4])
        s.descramble, s.c_init = 1, 12345 + c
    return sb.lib().srsb200_decode_tb_batch(eng.handle, arr, n_tb, 2)
run_scr(); run_scr()
eng.profile(True); eng.profile_read(); run_scr(); p = eng.profile_read(); eng.profile(False)
out["rm_rx_kernel_with_descrambling"] = {"ms": p["rm"][0], "GBps": out["rm_rx_kernel"]["algorithmic_MB"] / p["rm"][0] / 1e3}
print(json.dumps(out))
