import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, oracle_lib as ol, vecgen, srsran_4g_b200 as sb
o = ol.oracle(); eng = sb.Engine(0)
tbs, G, Qm, eb = 40, 300, 2, 0.5
tb = sb.TransportBlock(tbs); st = None
for tx, rv in enumerate((0, 2, 3, 1)):
    _, e = vecgen.make_tb(tbs, G, Qm, rv, eb, 77, scale=100)
    res_o = o.decode_tb(tbs, Qm, rv, e, 6, st); st = res_o["state"]
    ret = eng.decode_tb(tb, Qm, rv, e, 6)
    K = tb.seg["K1"]; L = 3*K+12
    print("tx", tx, "rv", rv, "ret", ret, res_o["ret"], "noi", tb.cb_noi[:1], res_o["cb_noi"][:1], "buf eq", (tb.buffer_f[0,:L]==st["buffer_f"][0,:L]).all(),
          "data eq", (tb.data[:K//8]==res_o["data"][:K//8]).all(), "cb_crc", tb.cb_crc[:1], st["cb_crc"][:1])
    # direct batch decode of the soft buffer with CRC24A
    out, noi, ok = eng.tdec_batch(K, st["buffer_f"][0:1,:L].copy(), 6, early_stop=True, crc_kind=sb.CRC_24A)
    hard = o.tdec_trace(K, st["buffer_f"][0,:L].copy(), 6)
    print("   tdec_batch noi", noi, "ok", ok, "oracle crcs", [o.crc_bytes(ol.CRC24A, hard[i], K) for i in range(6)], "match", [(out[0]==hard[i]).all() for i in range(6)])
