#!/usr/bin/env python3
"""How much does overlapping consecutive steps buy? Two engines (own streams + workspaces) decode the bench batch
concurrently on one GPU; aggregate throughput vs one engine alone."""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import srsran_4g_b200 as sb
from srsran_4g_b200 import synth
K, n_cb = 6144, 16384
dev = torch.device("cuda", 0)
bits, llr = synth.make_llr_batch(K, n_cb, 1.5, 1000, 100, n_distinct=256, device=dev)
def mk():
    eng = sb.Engine(0)
    d_out = torch.zeros((n_cb, K // 8), dtype=torch.uint8, device=dev); d_noi = torch.zeros(n_cb, dtype=torch.uint8, device=dev); d_ok = torch.zeros(n_cb, dtype=torch.uint8, device=dev)
    plan = eng.plan_uniform(n_cb, K, sb.CRC_24B)
    def step():
        eng.run_plan_dev(plan, llr.data_ptr(), 8, 2, True, d_out.data_ptr(), d_noi.data_ptr(), d_ok.data_ptr())
    return eng, step, (d_out, d_noi, d_ok)
for nengines in (1, 2, 3):
    es = [mk() for _ in range(nengines)]
    for e, s, _ in es:
        for _ in range(3): s()
        e.sync()
    steps = 24
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        es[i % nengines][1]()
    for e, _, _ in es: e.sync()
    dt = time.perf_counter() - t0
    print("engines %d: %.3f ms/step, %.1f Gbit/s" % (nengines, dt / steps * 1e3, steps * n_cb * K / dt / 1e9))
    for e, _, _ in es: e.close()
