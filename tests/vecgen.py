"""Seeded synthetic code blocks / transport blocks for the parity tests (encoder = the oracle's, which is itself pinned
to the reference encoder and to its K=504 known-answer vector)."""
import numpy as np

import oracle_lib as ol


def bits_to_bytes(bits):
    return np.packbits(np.asarray(bits, np.uint8))


def bytes_to_bits(b, n=None):
    bits = np.unpackbits(np.asarray(b, np.uint8))
    return bits if n is None else bits[:n]


def cb_payload(K, rng, with_crc=True, poly=ol.CRC24B):
    """K bits: K-24 random + CRC24 (so that the early-stop CRC can pass)"""
    o = ol.oracle()
    if not with_crc:
        return rng.integers(0, 2, K).astype(np.uint8)
    p = rng.integers(0, 2, K - 24).astype(np.uint8)
    crc = o.crc_bits(poly, p)
    return np.concatenate([p, np.array([(crc >> (23 - i)) & 1 for i in range(24)], np.uint8)])


def sigma_for(ebn0_db, rate=1.0 / 3.0):
    """true Eb/N0 convention: sigma^2 = 1 / (2 R Eb/N0) per unit-amplitude BPSK symbol"""
    return float(np.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0))))


def quantise(y, scale):
    v = np.trunc(scale * y)  # C cast: toward zero (turbodecoder_test.c:254)
    return np.clip(v, -32768, 32767).astype(np.int16)


def awgn_llr(bits, ebn0_db, scale, rng, rate=1.0 / 3.0):
    """bit 1 -> +1, bit 0 -> -1 (positive LLR <=> bit 1: turbodecoder_gen.c:266)"""
    s = 2.0 * np.asarray(bits, np.float64) - 1.0
    y = s + sigma_for(ebn0_db, rate) * rng.standard_normal(len(s))
    return quantise(y, scale)


def make_cb(K, ebn0_db, seed, scale=100, with_crc=True):
    """-> (info bits[K], llr int16[3K+12] natural layout)"""
    rng = np.random.default_rng(seed)
    bits = cb_payload(K, rng, with_crc)
    coded = ol.oracle().encode(bits)
    return bits, awgn_llr(coded, ebn0_db, scale, rng)


def make_cb_batch(K, n, ebn0_db, seed, scale=100, with_crc=True):
    L = 3 * K + 12
    llr = np.zeros((n, L), np.int16)
    bits = np.zeros((n, K), np.uint8)
    for i in range(n):
        bits[i], llr[i] = make_cb(K, ebn0_db, seed * 100003 + i, scale, with_crc)
    return bits, llr


def make_tb(tbs, G, Qm, rv, ebn0_db, seed, scale=100, payload=None):
    """Transport block through the oracle's encode chain: TB CRC24A, segmentation, CB CRC24B, turbo encode,
    rate matching with the decoder's E/rp convention (sch.c:397-407) -> (payload bytes, e_bits int16[G]).
    Standard TBS only (F == 0)."""
    o = ol.oracle()
    rng = np.random.default_rng(seed)
    ret, seg = o.cbsegm(tbs)
    assert ret == 0 and seg["F"] == 0
    if payload is None:
        payload = rng.integers(0, 2, tbs).astype(np.uint8)
    crc = o.crc_bits(ol.CRC24A, payload)
    tb = np.concatenate([payload, np.array([(crc >> (23 - i)) & 1 for i in range(24)], np.uint8)])
    Cn = seg["C"]
    e_tx = np.zeros(G, np.uint8)
    Gp = G // Qm
    gamma = Gp % Cn
    n_e = Qm * (Gp // Cn)
    rd = 0
    used = np.zeros(G, bool)
    for r in range(Cn):
        K = seg["K1"] if r < seg["C1"] else seg["K2"]
        rlen = K if Cn == 1 else K - 24
        cb = tb[rd:rd + rlen]
        rd += rlen
        if Cn > 1:
            c = o.crc_bits(ol.CRC24B, cb)
            cb = np.concatenate([cb, np.array([(c >> (23 - i)) & 1 for i in range(24)], np.uint8)])
        coded = o.encode(cb)
        E, rp = n_e, r * n_e
        if r > Cn - gamma:
            E = n_e + Qm
            rp = (Cn - gamma) * n_e + (r - (Cn - gamma)) * E
        e_tx[rp:rp + E] = o.rm_tx(coded, K, E, rv)
        used[rp:rp + E] = True
    rate = tbs / float(G)
    s = 2.0 * e_tx.astype(np.float64) - 1.0
    y = s + sigma_for(ebn0_db, rate) * rng.standard_normal(G)
    return bits_to_bytes(payload), quantise(y, scale)
