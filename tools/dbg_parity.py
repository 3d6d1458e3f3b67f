import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle_lib as ol, vecgen, srsran_4g_b200 as sb
o = ol.oracle(); eng = sb.Engine(0)
for K in (40, 64, 512, 6144):
    _, llr = vecgen.make_cb_batch(K, 2, 2.0, 3)
    for it in (1, 2, 3, 4):
        out, noi, ok = eng.tdec_batch(K, llr, it, early_stop=False, crc_kind=sb.CRC_24B)
        for c in range(2):
            exp = o.tdec_trace(K, llr[c], it)[it - 1]
            d = np.unpackbits(out[c]) != np.unpackbits(exp)
            print("K=%d it=%d cb=%d noi=%d ok=%d diffbits=%d first=%s" % (K, it, c, noi[c], ok[c], d.sum(), np.nonzero(d)[0][:12]))
