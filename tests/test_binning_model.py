"""CPU model of one claim the binning route's last sort pass relies on (csrc/radix_sort.cu, onesweep_tile (h), and
csrc/binning.cu, tile_offsets_fill_kernel): in the LAST pass of a complete LSD radix sort every 4096-pair tile is sorted
on the whole key, so the first slot of a key inside a digit run of a tile is that key's first position in the whole
output and may be written with a plain store, while the first slot of a digit run needs an atomicMin — and the result
does not depend on the order in which the thread blocks' stores and atomics land.  The model replays exactly those
operations in random order; the suffix minimum over the tiles then has to give the tile offsets."""
import numpy as np
import pytest

TILE = 4096


def _last_pass_operations(keys, shift):
    """keys: input of the last pass (sorted on the bits below `shift`).  -> (sorted output, [(op, key, position)])"""
    n = keys.size
    digit = (keys >> shift) & 255
    tiles = [np.arange(s, min(s + TILE, n)) for s in range(0, n, TILE)]
    # stable partition on the digit: all tiles' digit-0 runs in tile order, then digit 1, ...
    order = np.argsort(digit, kind="stable")
    pos_of = np.empty(n, dtype=np.int64)
    pos_of[order] = np.arange(n)
    ops = []
    for idx in tiles:
        local = idx[np.argsort(digit[idx], kind="stable")]   # the tile in shared memory: digit runs, stable inside
        k, d, dst = keys[local], digit[local], pos_of[local]
        for p in range(local.size):
            run_start = p == 0 or d[p - 1] != d[p]
            if run_start:
                ops.append(("min", int(k[p]), int(dst[p])))
            elif k[p - 1] != k[p]:
                ops.append(("store", int(k[p]), int(dst[p])))
    return keys[order], ops


@pytest.mark.parametrize("n,n_slots,seed", [(30_000, 3000, 0), (50_000, 40_000, 1), (9000, 70_000, 2), (4096, 300, 3), (1, 10, 4)])
def test_first_positions_do_not_depend_on_the_order_of_stores_and_atomics(n, n_slots, seed):
    rng = np.random.default_rng(seed)
    # many empty slots, a narrow band of busy ones and one very long list that spans several tiles of the pass —
    # like the (camera, tile) indices of a scene that covers part of the image
    keys = rng.integers(0, n_slots, size=n, dtype=np.int64)
    busy = rng.random(n) < 0.35
    keys[busy] = rng.integers(n_slots // 3, n_slots // 3 + max(n_slots // 50, 1), size=int(busy.sum()))
    keys[rng.random(n) < 0.3] = n_slots // 2
    bits = max(1, int(n_slots - 1).bit_length())
    passes = (bits + 7) // 8
    cur = keys
    for p in range(passes - 1):  # all passes but the last: plain stable LSD passes
        cur = cur[np.argsort((cur >> (8 * p)) & 255, kind="stable")]
    out, ops = _last_pass_operations(cur, 8 * (passes - 1))
    assert np.array_equal(out, np.sort(keys, kind="stable"))
    stores = [k for op, k, _ in ops if op == "store"]
    assert len(stores) == len(set(stores)), "at most one plain store per key"
    expected = np.full(n_slots, np.iinfo(np.int64).max)
    first = np.searchsorted(out, np.arange(n_slots), side="left")
    present = np.zeros(n_slots, dtype=bool)
    present[np.unique(keys)] = True
    expected[present] = first[present]
    for trial in range(4):
        order = rng.permutation(len(ops))
        got = np.full(n_slots, np.iinfo(np.int64).max)
        for i in order:
            op, k, dst = ops[i]
            got[k] = min(got[k], dst) if op == "min" else dst
        assert np.array_equal(got, expected), f"trial {trial}"
    # tile_offsets_fill_kernel: offsets[t] = first position of the first non-empty slot at or behind t, n behind the last
    filled = np.minimum.accumulate(np.minimum(expected, n)[::-1])[::-1]
    assert np.array_equal(filled, first)


def test_emission_owner_formula():
    """isect_scan_emit_kernel: the owner of entry k among a warp's (compacted, non-empty) runs is
    #runs that start before the 32-entry window + #runs that start inside it at or before k, minus one —
    one ballot, one warp-wide OR and two popcounts in the kernel.  Checked against the plain expansion."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        cnt = rng.integers(0, 40, size=32) * (rng.random(32) < 0.7)
        if trial == 0:
            cnt = np.zeros(32, dtype=np.int64); cnt[31] = 5          # only the last lane has a run
        if trial == 1:
            cnt = np.full(32, 1)                                     # 32 runs of one entry
        if trial == 2:
            cnt = np.zeros(32, dtype=np.int64); cnt[0] = 8160        # one run far longer than a window
        nonempty = np.flatnonzero(cnt > 0)
        starts_all = np.cumsum(cnt) - cnt
        start = np.full(32, np.iinfo(np.int32).max, dtype=np.int64)  # what lane l holds after the compaction
        start[:nonempty.size] = starts_all[nonempty]
        total = int(cnt.sum())
        truth = np.repeat(np.arange(nonempty.size), cnt[nonempty])   # owner (compacted index) of every entry
        for k0 in range(0, total, 32):
            before = int((start < k0).sum())
            d = start - k0
            inside = 0
            for l in range(32):
                if 0 <= d[l] < 32:
                    inside |= 1 << int(d[l])
            for lane in range(min(32, total - k0)):
                le_mask = (2 << lane) - 1
                owner = before + bin(inside & le_mask).count("1") - 1
                assert owner == truth[k0 + lane], (trial, k0, lane)


def _lookback_block(status, blk, total, out):
    """Generator model of chain_lookback (csrc/binning.cu) for one thread block: yields at every memory operation so a
    scheduler can interleave the blocks.  status[b] = (flag, value): flag 0 nothing, 1 aggregate, 2 inclusive prefix."""
    status[blk] = (2 if blk == 0 else 1, total)
    yield
    excl = 0
    if blk > 0:
        j = blk - 1
        done = False
        while not done:
            words = []
            for q in range(2):                       # a lane polls blk-1-lane and blk-33-lane: 64 loads, each at its own time
                chunk = []
                for lane in range(32):
                    jj = j - 32 * q - lane
                    chunk.append(status[jj] if jj >= 0 else (2, 0))   # before block 0: an empty prefix
                    yield
                words.append(chunk)
            for q in range(2):
                if done:
                    break
                flags = [w[0] for w in words[q]]
                first_nr = flags.index(0) if 0 in flags else 32
                first_pf = flags.index(2) if 2 in flags else 32
                take = first_pf + 1 if first_pf < first_nr else first_nr
                excl += sum(w[1] for w in words[q][:take])
                j -= take
                if first_pf < first_nr:
                    done = True
                elif take < 32:
                    break
        status[blk] = (2, excl + total)
        yield
    out[blk] = excl


@pytest.mark.parametrize("n_blocks,resident,seed", [(1, 1, 0), (40, 40, 1), (300, 37, 2), (700, 148, 3), (200, 1, 4)])
def test_decoupled_lookback_under_random_interleaving(n_blocks, resident, seed):
    """Blocks start in ticket order, at most `resident` at a time, and advance one memory operation at a time in a
    random order: every block must still come out with the exact exclusive prefix of the blocks before it."""
    rng = np.random.default_rng(seed)
    totals = rng.integers(0, 10_000, size=n_blocks).tolist()
    status = {b: (0, 0) for b in range(n_blocks)}
    out = {}
    nxt = 0
    running = []
    steps = 0
    while len(out) < n_blocks:
        while nxt < n_blocks and len(running) < resident:      # a block slot frees up: the next ticket starts
            running.append(_lookback_block(status, nxt, totals[nxt], out))
            nxt += 1
        i = int(rng.integers(0, len(running)))
        try:
            next(running[i])
        except StopIteration:
            running.pop(i)
        steps += 1
        assert steps < 5_000_000, "the chain does not make progress"
    prefix = np.cumsum([0] + totals[:-1])
    assert [out[b] for b in range(n_blocks)] == prefix.tolist()
