#!/usr/bin/env python3
"""Per-kernel time of one large transport-block submission (device-resident soft buffers), DL and UL source."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import srsran_4g_b200 as sb
import vecgen
n_tb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tbs, G, Qm = 75376, 86400, 6
els = [vecgen.make_tb(tbs, G, Qm, 0, 6.0, 500 + c, scale=700)[1] for c in range(4)]
eng = sb.Engine(0)
eng.softbuffer_set_resident(True)
tbl = [sb.TransportBlock(tbs) for _ in range(n_tb)]
def run(ul):
    reqs = []
    for c, tb in enumerate(tbl):
        eng.softbuffer_reset(tb)
        reqs.append((tb, Qm, 0, els[c % 4], G // Qm, 12, G, (), 0, 0) if ul else (tb, Qm, 0, els[c % 4]))
    t0 = time.perf_counter()
    r = eng.ulsch_decode_batch(reqs, 8) if ul else eng.decode_tb_batch(reqs, 8)
    return time.perf_counter() - t0
for ul in (False, True):
    run(ul); run(ul)
    dt = min(run(ul) for _ in range(3))
    eng.profile(True); eng.profile_read(); run(ul); p = eng.profile_read(); eng.profile(False)
    print(json.dumps({"source": "ul q_bits" if ul else "dl e_bits", "n_tb": n_tb, "call_ms": dt * 1e3, "info_Gbit_s_call": n_tb * tbs / dt / 1e9,
                      "kernel_ms": {k: round(v[0], 3) for k, v in p.items() if v[1]}, "kernel_sum_ms": round(sum(v[0] for v in p.values()), 3)}))
