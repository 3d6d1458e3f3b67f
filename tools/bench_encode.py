#!/usr/bin/env python3
"""Transport-block encode throughput: srsb200_encode_tb_batch (device kernels by CUDA events + end-to-end with host
buffers) next to the reference's LUT encoder + rate matcher (oracle/_ref, one host core). Writes one JSON line.
usage: python tools/bench_encode.py [n_tb] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import srsran_4g_b200 as sb  # noqa: E402


def main():
    n_tb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tbs, Qm, G = 75376, 6, 86400      # 100 PRB, 64QAM, 13 code blocks of K = 6144
    rng = np.random.default_rng(0)
    reqs = [(tbs, Qm, 0, G, rng.integers(0, 256, tbs // 8, dtype=np.uint8)) for _ in range(n_tb)]
    eng = sb.Engine(0)
    for _ in range(3):
        ret, res = eng.encode_tb_batch(reqs)
        assert ret == 0
    eng.profile(True)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        ret, res = eng.encode_tb_batch(reqs)
    e2e_s = (time.perf_counter() - t0) / reps
    prof = eng.profile_read()
    eng.profile(False)
    kern_ms = (prof["tbenc"][0] + prof["tbcrc"][0]) / reps
    bits = n_tb * tbs
    out = dict(workload=f"{n_tb} TB x tbs {tbs} (13 CB of K=6144), Qm {Qm}, G {G}, rv 0", n_tb=n_tb,
               kernel_ms=kern_ms, kernel_gbps=bits / kern_ms / 1e6, e2e_ms=e2e_s * 1e3, e2e_gbps=bits / e2e_s / 1e9,
               kernel_detail=prof)
    try:
        import oracle_lib as ol
        ref = ol.ref()
        if ref is not None:
            m = min(n_tb, 64)
            t0 = time.perf_counter()
            for r in reqs[:m]:
                rr, e_ref = ref.encode_tb(*r)
            dt = time.perf_counter() - t0
            out["cpu_reference"] = dict(gbps=m * tbs / dt / 1e9, cores=1, sample=f"{m} TB", kind="reference")
            assert np.array_equal(e_ref[:G // 8], res[m - 1][1][:G // 8]), "device e-bits differ from the reference's"
            out["checked_against_reference"] = True
    except ImportError:
        pass
    line = json.dumps(out)
    print(line)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
