"""Host-side model of the streaming cascade's schedule (sift_project_b200/csrc/stream.cuh: StreamGeom::DEPTH,
the StreamSched block of k_stream, the running ring offsets of stream_level, StreamFetch).

The kernel itself is checked bit-for-bit on the GPU (tests/test_gpu_parity.py, scratch/stream_test.cu); this
test pins the INVARIANTS its comments claim, for every band position and many image heights, without a GPU:

  * a level only ever reads rows of the previous level that were completed in an EARLIER step
    (one CTA barrier per step is enough), or fetched by cp.async at least PF steps earlier;
  * a ring slot is never overwritten while a row it holds can still be read (ring depths
    2R + 3K - 1 and K*PF + 2R + 3K - 1);
  * reading outside the image reads the clamped edge row (the reference's clamp-to-edge, image.cpp:170-184);
  * every row of the band is completed exactly once by every level, in order.
"""
import itertools

import pytest


def schedule(h, y0, y1, radii, K):
    """The StreamSched block of k_stream."""
    NL = len(radii)
    R = [0] + list(radii)
    f = [0] * (NL + 1)
    e = [0] * (NL + 1)
    i0 = [0] * (NL + 1)
    f[NL], e[NL] = y0, y1 - 1
    for l in range(NL, 0, -1):
        i0[l] = f[l] - R[l]
        f[l - 1] = max(i0[l], 0)
        e[l - 1] = min(e[l] + R[l], h - 1)
    T = [0] * (NL + 2)
    Tend = [0] * (NL + 1)
    for l in range(1, NL + 1):
        Tend[l] = T[l] + (e[l] + R[l] - i0[l]) // K
        if l < NL:
            T[l + 1] = T[l] + 1 + (f[l] + K - 1 + R[l] - i0[l]) // K
    return dict(R=R, f=f, e=e, i0=i0, T=T, Tend=Tend, steps=Tend[NL] + 1, r0=f[0], rlast=e[0])


def simulate(h, y0, y1, radii, K, PF):
    sc = schedule(h, y0, y1, radii, K)
    NL, R = len(radii), sc["R"]
    depth = [K * PF + 2 * R[1] + 3 * K - 1] + [2 * R[l + 1] + 3 * K - 1 for l in range(1, NL)]
    # ring[l][slot] = (row held, step at which it became readable); level 0 = fetched input rows
    ring = [dict() for _ in range(NL)]
    done = [[] for _ in range(NL + 1)]          # rows completed per level, in order
    next_fetch = sc["r0"]
    reads_this_step = []
    reads_prev_step = []   # cp.async of step t is issued BEFORE barrier t: warps may still be reading step t-1's rows

    def fetch_rows(step_ready):
        nonlocal next_fetch
        for _ in range(K):
            if next_fetch <= sc["rlast"]:
                write(0, next_fetch, step_ready)
                next_fetch += 1

    def write(l, row, step_ready):
        slot = row % depth[l]
        old = ring[l].get(slot)
        if old is not None:
            # the row being overwritten must not be read in this step or later (for fetched rows: nor in the
            # previous step, whose readers are not separated from this cp.async by a barrier)
            live = reads_this_step + (reads_prev_step if l == 0 else [])
            assert all(not (ll == l and rr == old[0]) for ll, rr in live), (l, row, old)
        ring[l][slot] = (row, step_ready)

    def read(l, row, t):
        slot = row % depth[l]
        held = ring[l].get(slot)
        assert held is not None and held[0] == row, ("row not in its ring slot", l, row, held, t)
        assert held[1] <= t, ("row read before it is visible", l, row, held, t)
        reads_this_step.append((l, row))

    # prologue: PF commit groups; group g is complete (visible after the barrier) at step g
    for g in range(PF):
        fetch_rows(step_ready=g)
    state = {l: dict(i=sc["i0"][l]) for l in range(1, NL + 1)}
    for t in range(sc["steps"]):
        reads_prev_step[:] = reads_this_step
        reads_this_step.clear()
        fetch_rows(step_ready=t + PF)
        # all levels run concurrently between two barriers: collect reads first, then writes
        writes = []
        for l in range(1, NL + 1):
            if t < sc["T"][l] or t > sc["Tend"][l]:
                continue
            i = state[l]["i"]
            for k in range(K):
                # window row: the virtual row clamped to the image.  Virtual rows beyond e[l] + R only feed rows
                # nobody needs (the last, partial step of a level), so what they read does not matter.
                if i + k <= sc["e"][l] + R[l]:
                    ri = min(max(i + k, 0), h - 1)
                    assert sc["f"][l - 1] <= ri <= sc["e"][l - 1]
                    read(l - 1, ri, t)
                y = i + k - R[l]
                if sc["f"][l] <= y <= sc["e"][l]:
                    if y0 <= y < y1:
                        read(l - 1, y, t)          # DoG centre
                    done[l].append(y)
                    if l < NL:
                        writes.append((l, y))
            state[l]["i"] = i + K
        for l, y in writes:
            write(l, y, step_ready=t + 1)          # visible after the next barrier
    for l in range(1, NL + 1):
        assert done[l] == list(range(sc["f"][l], sc["e"][l] + 1)), (l, done[l][:5], sc["f"][l], sc["e"][l])
    return sc


GEOMS = [((4, 5, 6), 1, 12), ((8, 10), 2, 6), ((4, 5, 6), 2, 6), ((8, 10), 1, 12)]


@pytest.mark.parametrize("radii,K,PF", GEOMS)
def test_every_band_of_small_images(radii, K, PF):
    for h in list(range(1, 40)) + [47, 48, 49, 63, 64, 65, 97, 131]:
        bands = {(0, h)}
        for n in (2, 3, 5):
            hs = -(-h // n)
            bands |= {(j * hs, min((j + 1) * hs, h)) for j in range(n) if j * hs < h}
        for y0, y1 in sorted(bands):
            simulate(h, y0, y1, radii, K, PF)


@pytest.mark.parametrize("radii,K,PF", GEOMS)
def test_bands_of_a_tall_image(radii, K, PF):
    h = 4320
    for y0, y1 in itertools.chain([(0, 480), (480, 960), (3840, 4320), (4319, 4320), (0, 1), (17, 18)],
                                  [(j * 393, min((j + 1) * 393, h)) for j in range(11)]):
        sc = simulate(h, y0, y1, radii, K, PF)
        # pipeline fill: every level needs 2R rows of input before its first row completes, so a band costs its
        # rows plus 2 sum(R) rows (32 of ~430 for the first kernel, 36 of ~360 for the second) and a step per level
        assert sc["steps"] <= -(-(y1 - y0 + 2 * sum(radii)) // K) + 2 * len(radii) + 1


def test_schedule_matches_the_documented_start_rule():
    """Level l+1 starts the step after level l completed the last row of level l+1's first step."""
    sc = schedule(1000, 300, 500, (4, 5, 6), 1)
    for l in (1, 2):
        first_needed = max(sc["i0"][l + 1], 0)
        step_completed = sc["T"][l] + (first_needed + sc["R"][l] - sc["i0"][l])
        assert sc["T"][l + 1] == step_completed + 1


def test_random_bands_property():
    """Property test (hypothesis): any image height, any band inside it, any of the shipped / tested geometries."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(1, 700), st.data())
    def run(h, data):
        y0 = data.draw(st.integers(0, h - 1))
        y1 = data.draw(st.integers(y0 + 1, h))
        radii, K, PF = data.draw(st.sampled_from(GEOMS))
        sc = simulate(h, y0, y1, radii, K, PF)
        assert sc["steps"] >= -(-(y1 - y0) // K)

    run()
