"""Pins the CPU oracle (oracle/sift_oracle.cpp) against golden vectors produced by the REAL
reference (tests/golden/make_golden.py), and -- when oracle/_ref is prebuilt -- against the real
reference itself.  No GPU needed."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import oracle as O

FIELDS = ("x", "y", "octave", "layer", "size", "pori")


def same_kps(a, b, desc):
    assert len(a) == len(b)
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f
    if desc:
        assert np.array_equal(a["desc"], b["desc"])


@pytest.fixture(scope="module")
def synth(golden_dir):
    return np.load(os.path.join(golden_dir, "synth_256x192.npz"))


@pytest.fixture(scope="module")
def config1(golden_dir):
    return np.load(os.path.join(golden_dir, "config1.npz"))


def test_keypoint_layout():
    assert O.KP_DTYPE.itemsize == 168  # sift.hh:15-23


def test_generator_d_is_reproducible(synth):
    assert np.array_equal(O.synth_image(192, 256, seed=42), synth["image"])


def test_sigmas_and_taps(synth):
    run = O.Run(O.port(), synth["image"], keep_pyramid=False)
    assert np.array_equal(run.sigmas(), synth["sigmas"])
    # radii 4 (initial blur, sigma 1.249), then 4,5,6,8,10 (image.cpp:226)
    taps = np.zeros(32)
    lens = [O.port().lib.oracle_gaussian_taps(float(s), taps.ctypes.data, 32)
            for s in [np.sqrt(1.6 ** 2 - 1)] + list(synth["sigmas"][1:])]
    assert lens == [5, 5, 6, 7, 9, 11]


def test_port_matches_reference_golden_synth(synth):
    run = O.Run(O.port(), synth["image"], keep_pyramid=True)
    assert run.octaves == int(synth["octaves"]) == 7
    assert np.array_equal(run.extrema().astype(np.int32), synth["extrema"])
    same_kps(run.keypoints(0), synth["raw"], False)
    same_kps(run.keypoints(1), synth["oriented"], False)
    same_kps(run.keypoints(2), synth["final"], True)
    assert np.array_equal(run.gaussian(1, 3), synth["g_o1_l3"])
    assert np.array_equal(run.dog(1, 2), synth["dog_o1_l2"])
    for o, l, is_dog, total, amax, centre in synth["plane_checks"]:
        p = run.dog(int(o), int(l)) if is_dog else run.gaussian(int(o), int(l))
        assert p.sum() == total and np.abs(p).max() == amax
        assert p[p.shape[0] // 2, p.shape[1] // 2] == centre


def test_port_config1_known_answers(config1, golden_dir):
    """image1 5441/1077/1296/1286, image2 5767/1173/1448/1430, 269 matches."""
    want = {"image1": (5441, 1077, 1296, 1286), "image2": (5767, 1173, 1448, 1430)}
    finals = []
    for name in ("image1", "image2"):
        px = np.asarray(Image.open(os.path.join(golden_dir, name + ".png")))
        run = O.Run(O.port(), px, keep_pyramid=False)
        got = (len(run.extrema()), len(run.keypoints(0)), len(run.keypoints(1)), len(run.keypoints(2)))
        assert got == want[name]
        same_kps(run.keypoints(2), config1[name + "_final"], True)
        same_kps(run.keypoints(0), config1[name + "_raw"], False)
        finals.append(run.keypoints(2))
    ia, ib, d = O.match(O.port(), finals[0]["desc"], finals[1]["desc"])
    assert len(ia) == 269
    assert np.array_equal(ia, config1["match_ia"]) and np.array_equal(ib, config1["match_ib"])
    assert np.array_equal(d, config1["match_dist"])


def test_port_match_kat(golden_dir):
    k = np.load(os.path.join(golden_dir, "match_kat.npz"))
    a, b = k["a"], k["b"]
    cases = {"full": (a, b), "b1": (a[:20], b[:1]), "b0": (a[:20], b[:0]), "a0": (a[:0], b),
             "b2": (a[:50], b[:2])}
    for tag, (x, y) in cases.items():
        ia, ib, d = O.match(O.port(), x, y)
        assert np.array_equal(ia, k[tag + "_ia"]), tag
        assert np.array_equal(ib, k[tag + "_ib"]), tag
        assert np.array_equal(d, k[tag + "_dist"]), tag
    assert len(k["b1_ia"]) == 20 and len(k["b0_ia"]) == 0  # |B|=1 always matches, |B|=0 never


def test_integer_ratio_test_is_equivalent(golden_dir):
    """sqrt(d1) < 0.75 sqrt(d2)  <=>  16 d1 < 9 d2 on the integer squared distances
    (the form the GPU matcher uses)."""
    k = np.load(os.path.join(golden_dir, "match_kat.npz"))
    a, b = k["a"].astype(np.int64), k["b"].astype(np.int64)
    d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d2, axis=1, kind="stable")
    best, second = order[:, 0], order[:, 1]
    r = np.arange(len(a))
    keep = 16 * d2[r, best] < 9 * d2[r, second]
    assert np.array_equal(np.nonzero(keep)[0], k["full_ia"])
    assert np.array_equal(best[keep], k["full_ib"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not prebuilt")
def test_port_equals_real_reference_fresh_input():
    img = O.synth_image(160, 208, seed=99)
    r, p = O.Run(O.ref(), img), O.Run(O.port(), img)
    assert r.octaves == p.octaves
    assert np.array_equal(r.extrema(), p.extrema())
    for s in (0, 1, 2):
        same_kps(r.keypoints(s), p.keypoints(s), s == 2)
    for o in range(r.octaves):
        for l in range(6):
            assert np.array_equal(r.gaussian(o, l), p.gaussian(o, l))
        for l in range(5):
            assert np.array_equal(r.dog(o, l), p.dog(o, l))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not prebuilt")
def test_rgb_and_undoubled_paths_equal_reference():
    rng = np.random.default_rng(5)
    g = O.synth_image(120, 150, seed=3).astype(np.float64)
    rgb = np.clip(np.stack([g, np.roll(g, 3, 1), np.roll(g, 5, 0)], -1) + rng.integers(-3, 4, (120, 150, 3)), 0, 255)
    for doubled in (True, False):
        r, p = O.Run(O.ref(), rgb, doubled), O.Run(O.port(), rgb, doubled)
        same_kps(r.keypoints(2), p.keypoints(2), True)
        assert len(r.keypoints(2)) > 10


@pytest.mark.skipif(not (O.have_ref() and O.have_ref(asshipped=True)), reason="oracle/_ref not prebuilt")
def test_copy_free_reference_build_is_bit_identical_to_as_shipped():
    img = O.synth_image(96, 128, seed=7)
    a, b = O.Run(O.ref(asshipped=True), img), O.Run(O.ref(), img)
    same_kps(a.keypoints(2), b.keypoints(2), True)
    assert np.array_equal(a.extrema(), b.extrema())


PARAM_SETS = [
    dict(init_sigma=1.3, contrast_threshold=0.03, eigen_ratio=6.0, peak_ratio=0.7, ori_sigma_factor=1.2,
         desc_scale_factor=2.5),
    dict(double_image_size=False, init_sigma=2.0, contrast_threshold=0.08, eigen_ratio=15.0, peak_ratio=0.9,
         ori_sigma_factor=1.8, desc_scale_factor=3.5),
    dict(intervals=2, contrast_threshold=0.05),
    dict(intervals=4, init_sigma=1.4, peak_ratio=0.75),
    dict(num_bins=18), dict(num_bins=72, peak_ratio=0.7), dict(window_size=5, intervals=4),
    dict(window_size=5, contrast_threshold=0.02),
]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not prebuilt")
@pytest.mark.parametrize("kw", PARAM_SETS)
def test_port_equals_real_reference_with_non_default_arguments(kw):
    """Every tunable argument of sift.hh:65-71 (intervals included) through both implementations."""
    img = O.synth_image(150, 200, seed=21)
    r = O.Run(O.ref(), img, params=O.Params(**kw))
    p = O.Run(O.port(), img, params=O.Params(**kw))
    assert r.octaves == p.octaves
    assert np.array_equal(r.sigmas(), p.sigmas())
    assert np.array_equal(r.extrema(), p.extrema())
    for s in (0, 1, 2):
        same_kps(r.keypoints(s), p.keypoints(s), s == 2)
    assert len(r.keypoints(2)) > 10
