"""Synthetic test-signal generation for bench.py / smoke: LTE turbo encoding (36.212 5.1.3.2) of random payloads with
CRC24B attached, vectorised across code blocks in numpy, then BPSK + AWGN + int16 quantisation (on the GPU when torch
has one). This is the encode direction of the path used only to make inputs; it mirrors
srsran_tcod_encode (lib/src/phy/fec/turbo/turbocoder.c:77-185) output order: s0 p0 p'0 s1 ... then 12 tail values."""
import numpy as np

# (pure numpy / torch: this module must not load libsrsran_b200.so - bench.py's reference arm generates its inputs here)

CRC24B_POLY = 0x1800063
CRC24A_POLY = 0x1864CFB


def qpp(K):
    """pi(i) = (f1 i + f2 i^2) mod K with the 36.212 Table 5.1.3-3 parameters held by the library's host tables"""
    assert K in _qpp_params(), "K=%d is not an LTE turbo block size" % K
    # recover f1, f2 through the published table in include/lte_qpp_params.h (parsed once)
    f1, f2 = _qpp_params()[K]
    i = np.arange(K, dtype=np.uint64)
    return ((f1 * i + f2 * i * i) % K).astype(np.int64)


_QPP = None


def _qpp_params():
    global _QPP
    if _QPP is None:
        import os
        import re
        hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lte_qpp_params.h")).read()
        _QPP = {int(k): (int(a), int(b)) for k, a, b in re.findall(r"\{(\d+),\s*(\d+),\s*(\d+)\}", hdr)}
    return _QPP


def crc24_attach(payload_bits, poly=CRC24B_POLY):
    """payload_bits [n, K-24] uint8 -> [n, K] with the 24-bit CRC (MSB first, zero init) appended; vectorised over n"""
    n, m = payload_bits.shape
    reg = np.zeros(n, np.uint32)
    p = np.uint32(poly & 0xFFFFFF)
    for i in range(m):
        fb = ((reg >> np.uint32(23)) & np.uint32(1)) ^ payload_bits[:, i].astype(np.uint32)
        reg = ((reg << np.uint32(1)) & np.uint32(0xFFFFFF)) ^ (fb * p)
    crc = np.stack([((reg >> np.uint32(23 - j)) & np.uint32(1)).astype(np.uint8) for j in range(24)], axis=1)
    return np.concatenate([payload_bits, crc], axis=1)


def _rsc(u):
    """constituent encoder g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3, with trellis termination.
    u [n, K] -> (parity [n, K], tail_sys [n, 3], tail_par [n, 3])"""
    n, K = u.shape
    d0 = np.zeros(n, np.uint8); d1 = np.zeros(n, np.uint8); d2 = np.zeros(n, np.uint8)
    z = np.zeros((n, K), np.uint8)
    for i in range(K):
        a = u[:, i] ^ d1 ^ d2
        z[:, i] = a ^ d0 ^ d2
        d2, d1, d0 = d1, d0, a
    ts = np.zeros((n, 3), np.uint8); tp = np.zeros((n, 3), np.uint8)
    for j in range(3):
        x = d1 ^ d2          # input that drives the register input to zero
        ts[:, j] = x
        tp[:, j] = d0 ^ d2   # a = 0
        d2, d1, d0 = d1, d0, np.zeros(n, np.uint8)
    return z, ts, tp


def turbo_encode(bits):
    """bits [n, K] uint8 -> coded [n, 3K+12] uint8 in the decoder's natural input order"""
    n, K = bits.shape
    pi = qpp(K)
    z1, ts1, tp1 = _rsc(bits)
    z2, ts2, tp2 = _rsc(bits[:, pi])
    out = np.zeros((n, 3 * K + 12), np.uint8)
    out[:, 0:3 * K:3] = bits
    out[:, 1:3 * K:3] = z1
    out[:, 2:3 * K:3] = z2
    for j in range(3):
        out[:, 3 * K + 2 * j] = ts1[:, j]
        out[:, 3 * K + 2 * j + 1] = tp1[:, j]
        out[:, 3 * K + 6 + 2 * j] = ts2[:, j]
        out[:, 3 * K + 6 + 2 * j + 1] = tp2[:, j]
    return out


def sigma_for(ebn0_db, rate=1.0 / 3.0):
    """true Eb/N0: sigma^2 = 1/(2 R Eb/N0) per unit-amplitude BPSK symbol"""
    return float(np.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0))))


def make_llr_batch(K, n, ebn0_db, seed, scale=100, n_distinct=256, device=None):
    """n code blocks of size K (payload K-24 random bits + CRC24B), BPSK (bit 1 -> +1), AWGN at the true Eb/N0,
    llr = (int16) trunc(scale * y). n_distinct different codewords are tiled over the batch, the noise is independent
    per block. Returns (bits [n_distinct, K] uint8, llr) with llr a torch tensor on `device` (or a numpy array when
    device is None)."""
    rng = np.random.default_rng(seed)
    nd = min(n, n_distinct)
    bits = crc24_attach(rng.integers(0, 2, (nd, K - 24)).astype(np.uint8))
    coded = turbo_encode(bits)
    sig = sigma_for(ebn0_db)
    if device is None:
        reps = (n + nd - 1) // nd
        s = np.tile(2.0 * coded.astype(np.float32) - 1.0, (reps, 1))[:n]
        y = s + sig * rng.standard_normal(s.shape, dtype=np.float32)
        return bits, np.clip(np.trunc(scale * y), -32768, 32767).astype(np.int16)
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sym = torch.from_numpy(2.0 * coded.astype(np.float32) - 1.0).to(device)
    llr = torch.empty((n, 3 * K + 12), dtype=torch.int16, device=device)
    for a in range(0, n, nd):
        b = min(n, a + nd)
        y = sym[: b - a] + sig * torch.randn((b - a, 3 * K + 12), generator=g, device=device)
        llr[a:b] = torch.clamp(torch.trunc(scale * y), -32768, 32767).to(torch.int16)
    return bits, llr
