"""smallest end-to-end case for compute-sanitizer: a few code blocks of three sizes + one transport block"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, oracle_lib as ol, vecgen, srsran_4g_b200 as sb
o = ol.oracle(); eng = sb.Engine(0)
for K, n in ((40, 3), (1024, 5), (6144, 3)):
    _, llr = vecgen.make_cb_batch(K, n, 1.5, 7)
    out, noi, ok = eng.tdec_batch(K, llr, 6, early_stop=True)
    _, oo, on, ook = o.tdec_batch(K, llr, 6, True)
    assert (out == oo).all() and (noi == on).all() and (ok == ook).all()
tbs, G, Qm = 6200, 9000, 4
_, e = vecgen.make_tb(tbs, G, Qm, 0, 1.5, 3)
tb = sb.TransportBlock(tbs)
assert eng.decode_tb(tb, Qm, 0, e, 6) == o.decode_tb(tbs, Qm, 0, e, 6)["ret"]
eng.close()
print("sanitizer case ok")
