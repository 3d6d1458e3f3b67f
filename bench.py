#!/usr/bin/env python
"""bench.py - decoded information throughput of the batched LTE turbo decoder (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): 16384 code blocks of K=6144, synthetic BPSK/AWGN LLRs at true Eb/N0 = 1.5 dB,
LLR scale 100, CRC24B early stop (min 2), at most 8 half-iterations (= 4 full turbo iterations; the reference API counts
half-iterations). One "step" = one pass of the hot path (de-multiplex -> turbo decode with fused
CRC early stop -> hard-bit emit) over the whole batch.

  value : whole-job decoded info Mbit/s with the LLRs already resident in HBM (CUDA events on the engine stream,
          max over ranks)
  e2e   : the same metric through the host-pointer C-ABI call (srsb200_tdec_batch) with pinned HOST buffers:
          H2D of the LLRs and D2H of bits / iteration counts / CRC flags inside the timed region
  roofline / roofline_int : the dominant kernel (tdec_group_kernel) against the HBM copy peak (algorithmic bytes) and
          against the measured packed-int16x2 issue peak (algorithmic op count) - SURVEY.md section 8(d)
  cpu_baseline : the reference's AVX2 windowed decoder compiled from the reference sources (oracle/_ref), all host
          cores, on a bounded sample of the same workload (N=1, rank 0 only)

--impl reference times that CPU implementation alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 6144
N_CB = 16384
EBN0_DB = 1.5
LLR_SCALE = 100
MAX_ITER = 8
MIN_ITER = 2
# 40.25 packed-s16x2 instructions per info bit per half-iteration (generic algorithm op count)
ALG_OPS_PER_BIT_HALFITER = 40.25
# measured by tools/microbench/int16x2_issue.cu on this pool's B200 (profiles/r01_int16x2_issue.txt): the two integer
# pipes together sustain 2 x 64 packed lanes/clk/SM -> 148 SM x 128 x 1.965 GHz
INT_PEAK_TOPS = 36.8
WORKLOAD = ("16384 code blocks x K=6144, true Eb/N0 1.5 dB BPSK/AWGN, int16 LLR scale 100, CRC24B early stop (min 2), "
            "max 8 half-iterations (= 4 full turbo iterations)")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU during the timed region (pynvml, ~10 ms period)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "applications_clocks_setting": 0x2, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def kernel_source_sha():
    """fingerprint of the source of the decode kernels the roofline is about (scan / job / extract / emit: turbo_kernels.cuh):
    ncu-derived figures (profiles/dram_traffic.json) are only printed for the kernels they were captured on"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "srsran_4g_b200", "csrc")
    for f in ("turbo_kernels.cuh",):
        h.update(f.encode())
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def bind_to_gpu_numa_node(torch, device_index):
    """One process per GPU on a multi-socket box: run on the cores next to this rank's GPU, so that the page-locked host buffers
    it allocates afterwards (first touch) sit in that socket's memory and the host-to-device copies of eight ranks do not all cross
    the socket interconnect. Reads the GPU's PCI address and sysfs; does nothing if either is unavailable (SRSB200_NO_NUMA_BIND=1:
    off). Returns a description for the bench line or None."""
    if os.environ.get("SRSB200_NO_NUMA_BIND"):
        return None
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= set(os.sched_getaffinity(0))
        if not cpus or cpus == set(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, cpus)
        node = None
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except Exception:  # noqa: BLE001
            pass
        return {"gpu": bdf, "numa_node": node, "cpus": len(cpus)}
    except Exception:  # noqa: BLE001
        return None


def cpu_decoder():
    """(library wrapper, kind, impl id, build description): oracle/_ref (the compiled reference, AVX2 windowed decoder, built
    with the reference's release flags) when it travelled with the repo, else the clean-room port. Loaded through
    tests/oracle_lib.py only - nothing of srsran_4g_b200's native code is touched on this leg."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    r, build = ol.ref_for_timing()
    if r is not None:
        return r, "reference", 5, build  # SRSRAN_TDEC_AVX_WINDOW / AUTO for K = 6144
    return ol.oracle(), "port", 1, "gcc -O2 (clean-room generic int16 port)"


CPU_CBS_PER_THREAD = 256


class CpuArm:
    """The reference's CPU implementation of the path on a bounded sample of the workload, all host threads, pinned. One
    'step' = cores x 64 code blocks; the decoder objects and the production sub-block input layout (what srsran_rm_turbo_rx_lut
    leaves in the soft buffer, sch.c:415) are prepared by every thread BEFORE its start barrier, as BASELINE.md section 3 says.
    Used by BOTH CPU legs (cpu_baseline of the b200 arm and --impl reference), so their figures are the same measurement."""

    def __init__(self, max_iter, seed):
        from srsran_4g_b200 import synth  # pure numpy on this path (no native library)
        self.lib, self.kind, self.impl, self.build = cpu_decoder()
        self.cores = len(os.sched_getaffinity(0))
        self.n = self.cores * CPU_CBS_PER_THREAD
        self.max_iter = max_iter
        _, self.llr = synth.make_llr_batch(K, self.n, EBN0_DB, seed, LLR_SCALE, n_distinct=64)
        self.noi = None

    def step(self):
        """-> seconds inside the library between the start barrier and the last thread finishing"""
        if self.kind == "reference":
            secs, _, self.noi, _ = self.lib.tdec_batch(K, self.llr, self.max_iter, True, nthreads=self.cores, impl=self.impl, pin=3)
        else:
            secs, _, self.noi, _ = self.lib.tdec_batch(K, self.llr, self.max_iter, True, nthreads=self.cores)
        return secs

    def describe(self, steps, total_s):
        name = ("srsRAN AVX2 windowed int16 decoder (turbodecoder_win.h, selected by srsran_tdec_init for K=6144) compiled from the "
                "reference tree with %s, input in the rate de-matcher's sub-block layout" % self.build) if self.kind == "reference" \
            else "clean-room generic int16 port (oracle/turbo_oracle.c)"
        return "%d code blocks (%d per thread) of the same workload per step x %d steps (%.1f s), %s, %d pthreads pinned" % (
            self.n, CPU_CBS_PER_THREAD, steps, total_s, name, self.cores)


def run_cpu(max_iter, min_seconds, seed, warmup=2):
    arm = CpuArm(max_iter, seed)
    for _ in range(warmup):
        arm.step()
    total_s, reps = 0.0, 0
    while total_s < min_seconds or reps < 3:
        total_s += arm.step()
        reps += 1
    return {"value": arm.n * K * reps / total_s / 1e6, "unit": "Mbit/s", "cores": arm.cores, "kind": arm.kind,
            "sample": arm.describe(reps, total_s), "mean_half_iterations": float(np.mean(arm.noi))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-cb", type=int, default=N_CB, help="code blocks per GPU (default: the BASELINE config)")
    ap.add_argument("--max-iter", type=int, default=MAX_ITER)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads (one engine each) calling the synchronous host-pointer API")
    ap.add_argument("--no-pipeline", action="store_true", help="one plan: every step waits for the previous one to finish completely")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--plans", type=int, default=2, help="plans (workspaces) used alternately by the device-resident leg; the engine pipelines consecutive submissions of different plans")
    ap.add_argument("--no-llr8", action="store_true", help="skip the 8-bit LLR leg (e2e_llr8)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    peaks, peak_src = measured_peaks()

    # ------------------------------------------------------------------ reference arm: the CPU implementation alone
    if args.impl == "reference":
        if rank != 0:
            return
        arm = CpuArm(args.max_iter, 4321)
        for _ in range(max(args.warmup, 3)):  # clocks and caches of 16 freshly woken cores need a few passes
            arm.step()
        dt = 0.0
        for _ in range(args.steps):
            dt += arm.step()
        val = arm.n * K * args.steps / dt / 1e6
        sample = arm.describe(args.steps, dt)
        print(json.dumps({
            "impl": "reference", "metric": "turbo-decoded info Mbit/s (K=6144, 4 iter)", "value": val, "unit": "Mbit/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "note": "CPU arm: each step decodes a bounded sample of the workload (%d code blocks), not the 16384 of the GPU arm; "
                    "throughput in Mbit/s is size-independent for independent code blocks" % arm.n,
            "cpu_baseline": {"value": val, "unit": "Mbit/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import srsran_4g_b200 as sb
    from srsran_4g_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n_cb = args.n_cb
    eng = sb.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    bits, llr = synth.make_llr_batch(K, n_cb, EBN0_DB, 1000 + rank, LLR_SCALE, n_distinct=256, device=dev)
    torch.cuda.synchronize()
    # two plans (each owns a workspace and its outputs) used alternately: consecutive steps are pipelined by the engine -
    # the latency-bound last half-iterations of step i overlap the first ones of step i+1 (include/srsran_b200.h)
    NPLAN = 1 if args.no_pipeline else max(1, args.plans)
    outs = [(torch.zeros((n_cb, K // 8), dtype=torch.uint8, device=dev), torch.zeros(n_cb, dtype=torch.uint8, device=dev),
             torch.zeros(n_cb, dtype=torch.uint8, device=dev)) for _ in range(NPLAN)]
    plans = [eng.plan_uniform(n_cb, K, sb.CRC_24B) for _ in range(NPLAN)]
    d_out, d_noi, d_ok = outs[0]
    step_no = [0]

    def step():
        i = step_no[0] % NPLAN
        step_no[0] += 1
        eng.run_plan_dev(plans[i], llr.data_ptr(), args.max_iter, MIN_ITER, True, outs[i][0].data_ptr(), outs[i][1].data_ptr(), outs[i][2].data_ptr())

    for _ in range(warmup):
        step()
    eng.sync()
    # correctness gate on this rank's batch (not timed): CRC-passing blocks must equal the transmitted payloads
    noi = d_noi.cpu().numpy()
    ok = d_ok.cpu().numpy()
    tx = np.packbits(bits, axis=1)
    got = d_out.cpu().numpy()
    idx = np.arange(n_cb) % len(bits)
    good = ok == 1
    if not (got[good] == tx[idx[good]]).all():
        raise SystemExit("decoded bits differ from the transmitted payload on CRC-passing blocks")
    # ... and a sample of this very batch against the oracle (the reference's generic int16 algorithm): bytes, half-iteration
    # counts and CRC verdicts of 64 blocks spread over the batch. The checker only - never timed, never on the product path.
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    samp = np.unique(np.linspace(0, n_cb - 1, 64).astype(np.int64))
    _, o_out, o_noi, o_ok = ol.oracle().tdec_batch(K, llr[torch.from_numpy(samp).to(dev)].cpu().numpy(), args.max_iter, True,
                                                   nthreads=max(1, len(os.sched_getaffinity(0))))
    if not ((o_out == got[samp]).all() and (o_noi == noi[samp]).all() and (o_ok == ok[samp]).all()):
        raise SystemExit("bench batch: GPU results differ from the oracle on the sampled code blocks")
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    eng.sync()
    l0 = eng.launch_count
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    eng.flush()  # the engine stream now waits for every submission in flight; ev1 closes the timed region after them
    ev1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches = eng.launch_count - l0
    ms = ev0.elapsed_time(ev1)
    for o_ in outs[1:]:  # every plan decoded the same batch: identical results or the pipelining is broken
        if not (torch.equal(o_[0], outs[0][0]) and torch.equal(o_[1], outs[0][1]) and torch.equal(o_[2], outs[0][2])):
            raise SystemExit("pipelined plans disagree")
    clocks = sampler.result()
    # per-kernel durations: CUDA events around every launch on the launching stream. The engine runs the decode as a
    # single chain of launches while profiling (no sub-batch overlap), so these are un-inflated kernel times; they are
    # taken right after the timed region on the same inputs.
    psteps = max(1, min(args.steps, 5))
    eng.profile(True)
    eng.profile_read()
    for _ in range(psteps):
        step()
    prof = eng.profile_read()
    eng.profile(False)
    ms_per_rank = [ms / args.steps]
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        ms_per_rank = [float(x.item()) / args.steps for x in allms]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bits_per_step = float(n_cb) * K * world
    value = bits_per_step * args.steps / (ms * 1e-3) / 1e6

    # ------------------------------------------------------------------ end to end through the host-pointer C ABI
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        L = sb.lib()
        h_llr = torch.empty((n_cb, 3 * K + 12), dtype=torch.int16, pin_memory=True)
        h_llr.copy_(llr)
        h_out = torch.empty((n_cb, K // 8), dtype=torch.uint8, pin_memory=True)
        h_noi = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)
        h_ok = torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)
        Ks = np.full(n_cb, K, np.uint32)
        kinds = np.full(n_cb, sb.CRC_24B, np.uint8)
        loff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(3 * K + 12))
        ooff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(K // 8))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)

        # Two host threads, each with its own engine and result buffers, call the synchronous API alternately - the way a
        # PHY with several worker threads uses it (srsRAN runs 3-4). The copy of one call's LLRs then overlaps the ~1.3 ms
        # decode tail of the other call; every call still moves its full input and output across PCIe.
        import threading
        T = max(1, args.e2e_threads)
        engs = [eng] + [sb.Engine(local_rank) for _ in range(T - 1)]
        obufs = [(h_out, h_noi, h_ok)] + [(torch.empty((n_cb, K // 8), dtype=torch.uint8, pin_memory=True), torch.empty(n_cb, dtype=torch.uint8, pin_memory=True),
                                          torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)) for _ in range(T - 1)]

        def e2e_step(t=0):
            ho, hn, hk = obufs[t]
            r = L.srsb200_tdec_batch(engs[t].handle, n_cb, vp(Ks), vp(kinds), C.c_void_p(h_llr.data_ptr()), vp(loff), n_cb * (3 * K + 12),
                                     args.max_iter, MIN_ITER, 1, C.c_void_p(ho.data_ptr()), vp(ooff), n_cb * (K // 8),
                                     C.c_void_p(hn.data_ptr()), C.c_void_p(hk.data_ptr()))
            if r != 0:
                raise SystemExit("srsb200_tdec_batch failed: %s" % L.srsb200_last_error().decode())

        for t in range(T):
            for _ in range(2):
                e2e_step(t)
            if not (obufs[t][1].numpy() == noi).all():
                raise SystemExit("host-path iteration counts differ from the device-resident path")
        e_steps = max(T, (max(3, min(args.steps, 10)) // T) * T)
        if dist is not None:
            dist.barrier()
        start = threading.Barrier(T + 1)
        errs = []

        def worker(t):
            try:
                start.wait()
                for _ in range(e_steps // T):
                    e2e_step(t)  # synchronous: results are in the host buffers on return
            except BaseException as ex:  # noqa: BLE001
                errs.append(ex)

        ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        for th in ths:
            th.start()
        start.wait()
        t0 = time.perf_counter()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        if errs:
            raise SystemExit("e2e worker failed: %r" % errs[0])
        for t in range(T):
            if not (obufs[t][1].numpy() == noi).all():
                raise SystemExit("host-path iteration counts differ from the device-resident path")
        for e_ in engs[1:]:
            e_.close()
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": bits_per_step * e_steps / dt / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": int(n_cb * (3 * K + 12) * 2),
               "d2h_bytes_per_step": int(n_cb * (K // 8) + 2 * n_cb), "steps": e_steps, "ms_per_step": dt / e_steps * 1e3, "host_threads": T,
               "api": "srsb200_tdec_batch (host pointers, pinned, synchronous) called from %d host thread(s) with one engine each; wall clock over all calls" % T}

    # ------------------------------------------------------------------ the same workload in the reference's 8-bit LLR mode
    # (q->llr_is_8bit: int8 LLRs, the windowed saturating decoders of turbodecoder_win.h reproduced bit for bit - SURVEY.md 8(f).3):
    # half the PCIe bytes per information bit. Same code blocks, LLR scale 12 so that +-127 is rarely hit; end to end through
    # srsb200_tdec_batch8 with pinned host buffers, copies inside the timed region; checked against the 8-bit oracle first.
    e2e8 = None
    if not args.no_e2e and not args.no_llr8:
        import ctypes as C
        L = sb.lib()
        bits8, llr8_16 = synth.make_llr_batch(K, n_cb, EBN0_DB, 1000 + rank, 12, n_distinct=256, device=dev)
        h_llr8 = torch.empty((n_cb, 3 * K + 12), dtype=torch.int8, pin_memory=True)
        h_llr8.copy_(llr8_16.clamp(-127, 127).to(torch.int8))
        del llr8_16
        Ks = np.full(n_cb, K, np.uint32)
        kinds = np.full(n_cb, sb.CRC_24B, np.uint8)
        loff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(3 * K + 12))
        ooff = (np.arange(n_cb, dtype=np.uint64) * np.uint64(K // 8))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        T = max(1, args.e2e_threads)
        engs8 = [sb.Engine(local_rank) for _ in range(T)]
        obufs8 = [(torch.empty((n_cb, K // 8), dtype=torch.uint8, pin_memory=True), torch.empty(n_cb, dtype=torch.uint8, pin_memory=True),
                   torch.empty(n_cb, dtype=torch.uint8, pin_memory=True)) for _ in range(T)]

        def e2e8_step(t=0):
            ho, hn, hk = obufs8[t]
            r = L.srsb200_tdec_batch8(engs8[t].handle, n_cb, vp(Ks), vp(kinds), C.c_void_p(h_llr8.data_ptr()), vp(loff), n_cb * (3 * K + 12),
                                      args.max_iter, MIN_ITER, 1, C.c_void_p(ho.data_ptr()), vp(ooff), n_cb * (K // 8),
                                      C.c_void_p(hn.data_ptr()), C.c_void_p(hk.data_ptr()))
            if r != 0:
                raise SystemExit("srsb200_tdec_batch8 failed: %s" % L.srsb200_last_error().decode())

        for t in range(T):
            for _ in range(2):
                e2e8_step(t)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        samp = np.unique(np.linspace(0, n_cb - 1, 48).astype(np.int64))
        _, o_out, o_noi, o_ok = ol.oracle().tdec8_batch(K, h_llr8.numpy()[samp], args.max_iter, True, nthreads=max(1, len(os.sched_getaffinity(0))))
        for t in range(T):
            if not ((o_out == obufs8[t][0].numpy()[samp]).all() and (o_noi == obufs8[t][1].numpy()[samp]).all() and (o_ok == obufs8[t][2].numpy()[samp]).all()):
                raise SystemExit("8-bit mode: GPU results differ from the 8-bit oracle on the sampled code blocks")
        e_steps = max(T, (max(3, min(args.steps, 10)) // T) * T)
        if dist is not None:
            dist.barrier()
        start8 = threading.Barrier(T + 1)
        errs8 = []

        def worker8(t):
            try:
                start8.wait()
                for _ in range(e_steps // T):
                    e2e8_step(t)
            except BaseException as ex:  # noqa: BLE001
                errs8.append(ex)

        ths = [threading.Thread(target=worker8, args=(t,)) for t in range(T)]
        for th in ths:
            th.start()
        start8.wait()
        t0 = time.perf_counter()
        for th in ths:
            th.join()
        dt8 = time.perf_counter() - t0
        if errs8:
            raise SystemExit("e2e8 worker failed: %r" % errs8[0])
        # kernel time of one call (events around every launch on the launching stream)
        engs8[0].profile(True)
        engs8[0].profile_read()
        e2e8_step(0)
        prof8 = engs8[0].profile_read()
        engs8[0].profile(False)
        noi8, ok8 = obufs8[0][1].numpy().copy(), obufs8[0][2].numpy().copy()
        for e_ in engs8:
            e_.close()
        if dist is not None:
            t = torch.tensor([dt8], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt8 = float(t.item())
        e2e8 = {"value": bits_per_step * e_steps / dt8 / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": int(n_cb * (3 * K + 12)),
                "d2h_bytes_per_step": int(n_cb * (K // 8) + 2 * n_cb), "steps": e_steps, "ms_per_step": dt8 / e_steps * 1e3, "host_threads": T,
                "kernel_ms_per_step": prof8["decode"][0] + prof8["extract"][0], "kernel_Mbit_s": float(n_cb) * K / ((prof8["decode"][0] + prof8["extract"][0]) * 1e-3) / 1e6,
                "mean_half_iterations": float(noi8.mean()), "crc_ok_fraction": float(ok8.mean()),
                "api": "srsb200_tdec_batch8 (int8 LLRs, scale 12; host pointers, pinned, synchronous) from %d host thread(s); the reference's windowed "
                       "saturating int8 decoder (32 windows at K=6144) reproduced bit for bit - a weaker decoder than the int16 one, hence more half-iterations" % T}

    # ------------------------------------------------------------------ transport blocks end to end (secondary: BASELINE config 5)
    # 64 cells x one 100-PRB 64QAM uplink transport block (TBS 75376 = 13 code blocks of K=5824, G = 86400 e-bits) per subframe
    # through srsb200_decode_tb_batch: e-bits from host memory, rate de-matching + HARQ soft buffers resident on the device (reset
    # every subframe: new data), decode, TB CRC, payload bytes back - 2.3 bytes over PCIe per information bit. Four host threads
    # with one engine each, the way PHY workers call it. Vectors from the engine's own encoder + AWGN; every payload is checked.
    e2e_tb = None
    if not args.no_e2e:
        import ctypes as C
        from srsran_4g_b200.binding import _TbStruct
        L = sb.lib()
        tbs_t, G_t, Qm_t, cells, T = 75376, 86400, 6, 64, 4
        rng_t = np.random.default_rng(77)  # the same vectors on every rank: the payload check below is then the same at any N
        enc = sb.Engine(local_rank)
        payloads = [rng_t.integers(0, 256, tbs_t // 8, dtype=np.uint8) for _ in range(4)]
        e_llr = []
        for pl in payloads:
            r_, eb_ = enc.encode_tb(tbs_t, Qm_t, 0, G_t, pl)
            if r_ != 0:
                raise SystemExit("encode_tb failed: %s" % L.srsb200_last_error().decode())
            sym = 2.0 * np.unpackbits(eb_)[:G_t].astype(np.float64) - 1.0
            e_llr.append(np.clip(np.trunc(40.0 * (sym + 0.40 * rng_t.standard_normal(G_t))), -32768, 32767).astype(np.int16))
        enc.close()
        engs_t = [sb.Engine(local_rank) for _ in range(T)]
        sets = []
        for t in range(T):
            engs_t[t].softbuffer_set_resident(True)
            tbl = [sb.TransportBlock(tbs_t) for _ in range(cells)]
            arr = (_TbStruct * cells)()
            for c, (st_, tb) in enumerate(zip(arr, tbl)):
                tb.fill(st_, Qm_t, 0, e_llr[c % 4])
            sets.append((tbl, arr, [(tb._bf, tb.max_cb) for tb in tbl]))

        def tb_step(t):
            tbl, arr, rst = sets[t]
            h_ = engs_t[t].handle
            for bf, mc in rst:  # new data in every cell: srsran_softbuffer_rx_reset forwarded to the device mirror
                L.srsb200_softbuffer_reset(h_, bf, mc)
            for tb in tbl:
                tb.cb_crc.fill(0)
            if L.srsb200_decode_tb_batch(engs_t[t].handle, arr, cells, args.max_iter) != 0:
                raise SystemExit("srsb200_decode_tb_batch failed: %s" % L.srsb200_last_error().decode())

        for t in range(T):
            tb_step(t)
            tb_step(t)
            tbl, arr, _ = sets[t]
            for c, st_ in enumerate(arr):
                if st_.ret != 0 or not np.array_equal(tbl[c].data[:tbs_t // 8], payloads[c % 4]):
                    raise SystemExit("e2e_tb: transport block %d of thread %d did not decode to its payload" % (c, t))
        sub_per_thread = 25
        if dist is not None:
            dist.barrier()
        start_t = threading.Barrier(T + 1)
        errs_t = []

        def worker_t(t):
            try:
                start_t.wait()
                for _ in range(sub_per_thread):
                    tb_step(t)
            except BaseException as ex:  # noqa: BLE001
                errs_t.append(ex)

        ths = [threading.Thread(target=worker_t, args=(t,)) for t in range(T)]
        for th in ths:
            th.start()
        start_t.wait()
        t0 = time.perf_counter()
        for th in ths:
            th.join()
        dt_t = time.perf_counter() - t0
        if errs_t:
            raise SystemExit("e2e_tb worker failed: %r" % errs_t[0])
        noi_t = float(np.mean([sets[0][0][c].cb_noi[:13].mean() for c in range(cells)]))
        for e_ in engs_t:
            e_.close()
        if dist is not None:
            tt = torch.tensor([dt_t], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_t = float(tt.item())
        nsub = T * sub_per_thread
        e2e_tb = {"value": world * nsub * cells * tbs_t / dt_t / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": cells * G_t * 2, "d2h_bytes_per_step": cells * (tbs_t // 8 + 13 + 14),
                  "bytes_over_pcie_per_info_bit": (cells * G_t * 2 + cells * (tbs_t // 8 + 27)) / float(cells * tbs_t), "steps": nsub,
                  "ms_per_step": dt_t / nsub * 1e3, "host_threads": T, "mean_half_iterations": noi_t,
                  "workload": "step = one subframe of %d cells x TBS %d (13 code blocks of K=5824, 64QAM, G=%d), HARQ soft buffers resident on the device" % (cells, tbs_t, G_t),
                  "api": "srsb200_decode_tb_batch (rate de-matching + soft combining + decode + TB CRC) from %d host threads with one engine each; aggregate" % T}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # dominant kernel = job_kernel (window-parallel recompute + LLR); "launch" below = all its launches of one step
    per_launch_ms = prof["job"][0] / psteps
    half_iters = float(noi.astype(np.float64).sum())  # executed half-iterations of this rank's batch (per launch)
    alg_bytes = n_cb * (2 * (3 * K + 12) + K // 8 + 2)  # LLRs in once, hard bits + flags out once
    hbm_achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    alg_ops = ALG_OPS_PER_BIT_HALFITER * K * half_iters
    int_achieved = alg_ops / (per_launch_ms * 1e-3) / 1e12
    # ncu-measured DRAM bytes: only for the kernel sources they were captured on (stale numbers are not printed)
    traffic, traffic_total, traffic_note = None, None, "no ncu capture of these kernel sources under profiles/dram_traffic.json"
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("kernel_source_sha") == kernel_source_sha():
            traffic, traffic_total = tj.get("job_kernel_bytes_per_step"), tj.get("total_bytes_per_step")
            traffic_note = tj.get("source")
        else:
            traffic_note = "profiles/dram_traffic.json was captured on other kernel sources (sha %s, now %s): not printed" % (
                tj.get("kernel_source_sha"), kernel_source_sha())
    except Exception:
        pass
    step_ms = ms / args.steps
    out = {
        "metric": "turbo-decoded info Mbit/s (K=6144, 4 iter)", "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (n_cb == N_CB and args.max_iter == MAX_ITER) else "NON-DEFAULT: %d CB, max_iter %d" % (n_cb, args.max_iter),
                   "code_blocks_per_gpu": n_cb, "K": K, "max_half_iterations": args.max_iter, "sharding": "independent code-block batches per GPU, no collective",
                   "pipelining": ("%d plans used alternately: consecutive steps overlap on the GPU" % NPLAN) if NPLAN > 1 else "none",
                   "l2": "inputs (%.0f MB LLRs + %.0f MB streams per step) exceed the 126 MB L2" % (n_cb * (3 * K + 12) * 2 / 1e6, n_cb * 5 * (K + 32) * 2 / 1e6)},
        "mean_half_iterations": float(noi.mean()), "crc_ok_fraction": float(ok.mean()),
        "noi_hist": {str(int(k)): int(v) for k, v in zip(*np.unique(noi, return_counts=True))},
        "clocks": clocks, "gpu_launches": int(launches), "ms_per_step_per_rank": ms_per_rank,
        "kernel_ms_per_step": {k: v[0] / psteps for k, v in prof.items() if v[1]},
        "kernel_launches_per_step": {k: v[1] // psteps for k, v in prof.items() if v[1]},
        "roofline": {"bound": "hbm", "kernel": "job_kernel (all half-iteration launches of one step, timed without sub-batch overlap)", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": hbm_achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": "MEASURED_PEAKS.json (%s)" % peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": per_launch_ms,
                     "measured_dram_GBps": (traffic / (per_launch_ms * 1e-3) / 1e9) if traffic else None,
                     "measured_dram_frac": (traffic / (per_launch_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic else None,
                     "traffic_source": traffic_note,
                     "note": "achieved = compulsory bytes of the whole decode (6.13 B/info bit, SURVEY 8(d)) over the job kernels' time; traffic / measured_dram_* = what ncu saw the job kernels move (profiles/dram_traffic.json); roofline_int = the integer-issue view"},
        "roofline_int": {"bound": "int16x2 issue (VIADD/VIMNMX/VIADDMNMX .16x2, two pipes)", "kernel": "job_kernel", "achieved": int_achieved,
                         "peak": INT_PEAK_TOPS, "unit": "T packed-instr/s", "frac": int_achieved / INT_PEAK_TOPS,
                         "algorithmic_ops_per_launch": alg_ops, "executed_half_iterations_per_launch": half_iters,
                         "peak_source": "tools/microbench/int16x2_issue.cu on this pool (profiles/r01_int16x2_issue.txt)"},
    }
    # the same two denominators over the WHOLE step as the driver times it (every kernel of the decode, pipelined), so that
    # work a kernel split moves into another kernel cannot hide
    out["roofline_step"] = {
        "int": {"achieved": alg_ops / (step_ms * 1e-3) / 1e12, "peak": INT_PEAK_TOPS, "unit": "T packed-instr/s",
                "frac": alg_ops / (step_ms * 1e-3) / 1e12 / INT_PEAK_TOPS},
        "hbm": {"achieved": alg_bytes / (step_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "measured_dram_bytes_per_step": traffic_total,
        "measured_dram_frac": (traffic_total / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic_total else None,
        "ms_per_step": step_ms}
    if e2e is not None:
        out["e2e"] = e2e
    if numa is not None:
        out["config"]["host_binding"] = "each rank runs on the cores local to its GPU (%s: NUMA node %s, %d cores)" % (numa["gpu"], numa["numa_node"], numa["cpus"])
    if e2e8 is not None:
        out["e2e_llr8"] = e2e8
    if e2e_tb is not None:
        out["e2e_tb"] = e2e_tb
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = run_cpu(args.max_iter, 10.0, 4321)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
