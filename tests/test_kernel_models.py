"""Host-side models of two pieces of device logic whose correctness argument is combinatorial, not numerical
(the kernels themselves are checked byte for byte on the GPU, tests/test_gpu_parity.py):

  * the ticket order of the tail kernel (sift_project_b200/csrc/pyramid.cu: k_tail, launch_tail): CTAs draw work
    items from one counter; group g holds the first-kernel tiles of tail octave g, then the second-kernel tiles of
    octave g - 1; a first-kernel tile of octave o waits for ALL first-kernel tiles of octave o - 1, a second-kernel
    tile of octave o for all first-kernel tiles of octave o.  The claim: every item depends only on items with
    LOWER tickets, hence no deadlock for ANY number of resident CTAs and any interleaving -- no cooperative launch;
  * the row lookup of the descriptor kernel's flattened sample list (detect.cu: k_describe, nth_set_bit): the
    non-empty rows of a 32-row block are compacted, and in every 32-sample iteration the rows that start in
    (s0, s0 + 32] set one bit each (REDUX.OR); a lane's row is the row of sample s0 plus the bits below its
    position.  The claim: the same (row, column) for every sample as the plain enumeration.
"""
import random

import numpy as np
import pytest


# ------------------------------------------------------------------------------------------
# k_tail
# ------------------------------------------------------------------------------------------
def tail_items(tiles):
    """launch_tail's ticket table: [(kind, octave, tile)] in ticket order; kind 'a' = G0 -> G1..G3 (+ next base),
    'b' = G3 -> D3, D4."""
    n = len(tiles)
    items = []
    for g in range(n + 1):
        if g < n:
            items += [("a", g, t) for t in range(tiles[g])]
        if g >= 1:
            items += [("b", g - 1, t) for t in range(tiles[g - 1])]
    return items


def deps(item, tiles):
    kind, o, _ = item
    dep = o - 1 if kind == "a" else o   # the octave whose first-kernel tiles this item reads
    return [("a", dep, t) for t in range(tiles[dep])] if dep >= 0 else []


@pytest.mark.parametrize("tiles", [[135, 40, 12, 6, 2, 1], [148, 1], [3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3], [1, 1],
                                   [54, 15, 6, 2, 1, 1]])
def test_tail_ticket_order_only_looks_back(tiles):
    items = tail_items(tiles)
    ticket = {it: k for k, it in enumerate(items)}
    assert len(ticket) == len(items) == 2 * sum(tiles)          # every tile of both kernels exactly once
    for it in items:
        for d in deps(it, tiles):
            assert ticket[d] < ticket[it], (it, d)


@pytest.mark.parametrize("resident", [1, 2, 7, 148])
@pytest.mark.parametrize("seed", range(4))
def test_tail_makes_progress_with_any_number_of_resident_ctas(resident, seed):
    """Event simulation: `resident` CTAs, each either idle (draws the next ticket), waiting for its item's
    dependencies, or running for a random time; the scheduler picks CTAs in random order.  Must finish, and no
    item may start before its dependencies are done."""
    rng = random.Random(seed)
    tiles = [rng.randint(1, 40) for _ in range(rng.randint(2, 8))]
    items = tail_items(tiles)
    done_a = [0] * len(tiles)
    finished = set()
    next_ticket = 0
    ctas = [None] * resident          # None | [item, remaining_time or None while waiting]
    steps = 0
    while len(finished) < len(items):
        steps += 1
        assert steps < 200 * len(items), "no progress: deadlock"
        c = rng.randrange(resident)
        if ctas[c] is None:
            if next_ticket < len(items):
                ctas[c] = [items[next_ticket], None]
                next_ticket += 1
            continue
        item, left = ctas[c]
        kind, o, _ = item
        if left is None:              # polling the hand-over counter
            dep = o - 1 if kind == "a" else o
            if dep < 0 or done_a[dep] == tiles[dep]:
                for d in deps(item, tiles):
                    assert d in finished
                ctas[c][1] = rng.randint(1, 5)
            continue
        if left > 1:
            ctas[c][1] = left - 1
            continue
        finished.add(item)
        if kind == "a":
            done_a[o] += 1
        ctas[c] = None
    assert done_a == tiles


# ------------------------------------------------------------------------------------------
# k_describe: row lookup
# ------------------------------------------------------------------------------------------
def nth_set_bit(m, n):
    """detect.cu: position of the (n + 1)-th set bit of the 32-bit mask m, 32 if there are fewer."""
    pos = 0
    w = 16
    while w >= 1:
        c = bin(m & ((1 << w) - 1)).count("1")
        if n >= c:
            n -= c
            pos += w
            m >>= w
        w >>= 1
    return pos if (m & 1) and n == 0 else 32


def lookup_by_mask(lo, cnt):
    """What the 32 lanes compute for one 32-row block: [(row slot, column)] per sample, in sample order."""
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)      # exclusive scan
    total = int(np.sum(cnt))
    nonempty = sum(1 << r for r in range(32) if cnt[r] > 0)
    c_off, c_row = [], []
    for lane in range(32):                                                  # compaction by shuffles
        src = nth_set_bit(nonempty, lane)
        c_off.append(int(off[src & 31]) if src < 32 else 0x7FFFFFFF)
        c_row.append(int(lo[src & 31]) * 32 + (src & 31))
    out = []
    kb = 0
    for s0 in range(0, total, 32):
        starts = 0
        for lane in range(32):
            p = (c_off[lane] - s0 - 1) & 0xFFFFFFFF
            if p < 32:
                starts |= 1 << p
        for lane in range(32):
            k = kb + bin(starts & ((1 << lane) - 1)).count("1")
            s = s0 + lane
            if s >= total:
                continue
            out.append((c_row[k] & 31, (c_row[k] >> 5) + (s - c_off[k])))
        kb += bin(starts).count("1")
    return out


def test_nth_set_bit():
    rng = random.Random(1)
    for _ in range(2000):
        m = rng.getrandbits(32) & rng.getrandbits(32) if rng.random() < 0.5 else rng.getrandbits(32)
        bits = [b for b in range(32) if m >> b & 1]
        for n in range(33):
            assert nth_set_bit(m, n) == (bits[n] if n < len(bits) else 32)


@pytest.mark.parametrize("seed", range(40))
def test_descriptor_row_lookup_equals_plain_enumeration(seed):
    rng = np.random.default_rng(seed)
    style = seed % 4
    if style == 0:      # typical window: 30-100 samples per row
        cnt = rng.integers(30, 100, 32)
    elif style == 1:    # narrow rows, several rows inside one 32-sample iteration, empty rows in between
        cnt = rng.integers(0, 6, 32)
    elif style == 2:    # mostly empty (the last block of a window, clipped corners)
        cnt = np.where(rng.random(32) < 0.2, rng.integers(1, 300, 32), 0)
    else:               # rows that are exact multiples of 32 (boundaries fall on iteration starts)
        cnt = rng.integers(0, 4, 32) * 32
    lo = rng.integers(-2000, 2000, 32)
    want = [(r, int(lo[r]) + j) for r in range(32) for j in range(int(cnt[r]))]
    assert lookup_by_mask(lo, cnt) == want
    if seed == 0:
        assert lookup_by_mask(lo, np.zeros(32, np.int64)) == []
