"""Multi-GPU sharding of independent code-block batches: code blocks never exchange data, so a batch is
split into contiguous ranges balanced by the number of trellis steps (sum of K) and every rank decodes its own range on
its own GPU - no collective on the data path. Only when a single consumer wants every result, the hard bits, iteration
counts and CRC flags are gathered with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_bounds(K_per_block, world):
    """contiguous ranges [lo, hi) per rank with sum(K) as equal as a contiguous split allows.
    K_per_block: int array of block sizes (or a scalar count n for equal-size blocks, via np.ones(n))."""
    K = np.asarray(K_per_block, np.int64)
    n = len(K)
    if n == 0:
        return [(0, 0)] * world
    c = np.concatenate([[0], np.cumsum(K)])
    total = c[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(c, target, side="left"))
        # choose the boundary whose prefix sum is closest to the target
        if i > 0 and abs(c[i - 1] - target) <= abs(c[min(i, n)] - target):
            i -= 1
        cuts.append(max(cuts[-1], min(i, n)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_results(out_bytes, noi, crc_ok, bounds, dist, device=None):
    """all-gather the per-rank results into full-batch arrays on every rank.
    out_bytes [n_r, K/8] uint8, noi [n_r], crc_ok [n_r] are this rank's numpy arrays (equal-K batch)."""
    import torch
    world = dist.get_world_size()
    nmax = max(hi - lo for lo, hi in bounds)
    kb = out_bytes.shape[1]
    pack = np.zeros((nmax, kb + 2), np.uint8)
    n_r = out_bytes.shape[0]
    pack[:n_r, :kb] = out_bytes
    pack[:n_r, kb] = noi
    pack[:n_r, kb + 1] = crc_ok
    t = torch.from_numpy(pack)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    rows = [p.cpu().numpy()[: hi - lo] for p, (lo, hi) in zip(parts, bounds)]
    full = np.concatenate(rows, axis=0)
    return full[:, :kb].copy(), full[:, kb].copy(), full[:, kb + 1].copy()
