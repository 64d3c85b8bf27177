"""TEST INFRASTRUCTURE ONLY -- numpy restatement of csrc/masks.cu (masks on device, counter-based RNG).

The reference draws its masks from host RNG streams (`torch.randperm`, `Categorical.sample`, `np.random.RandomState`:
cwm/models/masking.py:347-376, :100-132; utils.py:152-213); the bit-exact mirror of THOSE is the product's
`masking.py`, pinned by tests/golden/masks_ref.npz.  The device generators are an opt-in alternative whose masks depend
only on (seed, global sample index); what pins them is
  * the generator: Philox4x32-10 against the published Random123 known-answer vectors (`KAT`),
  * every selection step is integer arithmetic, restated here operation by operation, so kernel == oracle bit for bit,
  * the distributional contract of the reference functions they stand in for (counts per row, uniformity, energy
    proportionality), checked in tests/test_device_masks.py.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM_UNIFORM, STREAM_ENERGY, STREAM_RECT = 0, 1, 2

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values; returns four uint64 arrays of 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & np.uint64(0xffffffff) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xffffffff, int(k1) & 0xffffffff
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return c0, c1, c2, c3


def philox_words(seed, sample, stream, sub, count):
    """Words 0..count-1 of the stream (sample, stream, sub): word i = output[i % 4] of counter (i // 4, sample, stream, sub)."""
    i = np.arange(count, dtype=np.uint64)
    out = philox4x32_10(i >> np.uint64(2), sample, stream, sub, seed & 0xffffffff, (seed >> 32) & 0xffffffff)
    stacked = np.stack(out, 0)                                   # [4, count]
    return stacked[(i & np.uint64(3)).astype(np.int64), np.arange(count)]


def mask_uniform(seed, row0, rows, visible_frames, mask_frames, h, w, clump, n_visible_cells):
    gh, gw = h // clump, w // clump
    n = gh * gw
    out = np.ones((rows, visible_frames + mask_frames, h, w), dtype=np.uint8)
    out[:, :visible_frames] = 0
    for r in range(rows):
        for f in range(mask_frames):
            keys = (philox_words(seed, row0 + r, STREAM_UNIFORM, f, n) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            for cell in np.sort(keys)[:n_visible_cells] & np.uint64(0xffffffff):
                cy, cx = (int(cell) // gw) * clump, (int(cell) % gw) * clump
                out[r, visible_frames + f, cy:cy + clump, cx:cx + clump] = 0
    return out.reshape(rows, -1)


def energy_table(probs, eps):
    """probs fp32 [B, n] -> inclusive integer cumulative table uint64 [B, n] (same fp32 operations as the kernel)."""
    p = np.asarray(probs, dtype=np.float32)
    eps = np.float32(eps)
    mn = p.min(-1, keepdims=True)
    v = np.maximum((p - mn).astype(np.float32) + eps, np.float32(0)).astype(np.float32)
    mx = v.max(-1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = ((v / mx).astype(np.float32) * np.float32(16777216.0)).astype(np.float32)
    q = np.where(mx > 0, np.nan_to_num(q), 1.0).astype(np.uint64)
    return np.cumsum(q, -1, dtype=np.uint64)


def mask_energy_sample(table, h, w, clump, seed, sample0, S, points, visible_frames):
    B, n = table.shape
    gw = w // clump
    out = np.ones((B, S, visible_frames + 1, h, w), dtype=np.uint8)
    out[:, :, :visible_frames] = 0
    for b in range(B):
        total = int(table[b, -1])
        cum = [int(t) for t in table[b]]
        for s in range(S):
            o = philox4x32_10(np.arange(points, dtype=np.uint64), sample0 + s, STREAM_ENERGY, b, seed & 0xffffffff,
                              (seed >> 32) & 0xffffffff)
            for pnt in range(points):
                r = (int(o[0][pnt]) << 32) | int(o[1][pnt])
                target = (r * total) >> 64
                cell = next(i for i, c in enumerate(cum) if c > target)   # first cell whose inclusive sum exceeds it
                cy, cx = (cell // gw) * clump, (cell % gw) * clump
                out[b, s, visible_frames, cy:cy + clump, cx:cx + clump] = 0
    return out.reshape(B * S, -1)


def rectangularize(masks, row0, seed, target_masked=-1):
    m = np.array(masks, dtype=np.uint8).copy()
    rows, N = m.shape
    counts = (m != 0).sum(-1)
    target = int(counts.min()) if target_masked < 0 else int(target_masked)
    for r in range(rows):
        excess = int(counts[r]) - target
        if excess <= 0:
            continue
        words = philox_words(seed, row0 + r, STREAM_RECT, 0, N)
        keys = (words << np.uint64(32)) | np.arange(N, dtype=np.uint64)
        keys = np.where(m[r] != 0, keys, np.uint64(0xffffffffffffffff))
        for idx in np.sort(keys)[:excess] & np.uint64(0xffffffff):
            m[r, int(idx)] = 0
    return m
