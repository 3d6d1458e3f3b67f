"""8-bit mode micro-bench: n code blocks of K through srsb200_tdec_batch8, kernel times from the engine's event profiling.
  python tools/bench8.py [n_cb] [K] [max_iter]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsran_4g_b200 as sb  # noqa: E402
from srsran_4g_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
it = int(sys.argv[3]) if len(sys.argv) > 3 else 8
_, l16 = synth.make_llr_batch(K, n, 1.5, 7, 12, n_distinct=64)
l8 = np.clip(l16, -127, 127).astype(np.int8)
e = sb.Engine(0)
e.tdec_batch8(K, l8, it)
e.profile(True)
e.profile_read()
out, noi, ok = e.tdec_batch8(K, l8, it)
p = e.profile_read()
ms = p["decode"][0] + p["extract"][0]
print("n=%d K=%d: kernels %.3f ms (%d launches), %.1f Mbit/s, mean noi %.2f ok %.3f" % (n, K, ms, p["decode"][1], n * K / ms / 1e3, noi.mean(), ok.mean()))
