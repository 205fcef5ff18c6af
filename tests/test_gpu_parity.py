"""GPU parity tests: the CUDA path, called through the C ABI (ctypes binding), against the CPU
oracle on the same inputs and against the golden vectors the real reference produced.

Criteria (BASELINE.json north_star): keypoint recall / precision >= 99.5 % at 0.01 px and 1e-3
relative size; descriptors within 1 quantisation level; match index lists identical."""
import json
import os

import numpy as np
import pytest
from PIL import Image

import parity as P
from oracle import oracle as O

pytestmark = pytest.mark.gpu

S = pytest.importorskip("sift_project_b200")
REPORT = {}


@pytest.fixture(scope="module")
def ctx():
    c = S.SiftContext(1040, 768)
    yield c
    c.close()


@pytest.fixture(scope="module")
def synth(golden_dir):
    return np.load(os.path.join(golden_dir, "synth_256x192.npz"))


@pytest.fixture(scope="module")
def config1(golden_dir):
    return np.load(os.path.join(golden_dir, "config1.npz"))


@pytest.fixture(scope="module", autouse=True)
def _dump_report():
    yield
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.json", "w") as f:
        json.dump(REPORT, f, indent=1, default=float)
    print("\nPARITY REPORT", json.dumps(REPORT, default=float))


def test_pyramid_planes_match_oracle(ctx, synth):
    """Every Gaussian and DoG plane of every octave vs the FP64 oracle (image.cpp:156-238,
    sift.cpp:161-225).  FP32 storage: a few ulp of 255."""
    img = synth["image"]
    ctx.debug_options(keep_all_planes=True)   # the fused octave kernel keeps G4, G5 on chip otherwise
    ctx.detect(img)
    ctx.debug_options()
    run = O.Run(O.port(), img, keep_pyramid=True)
    st = ctx.stats()
    assert st["octaves"] == run.octaves == 7
    worst_g = worst_d = 0.0
    for o in range(run.octaves):
        assert ctx.plane_dims(o) == run.dims(o)
        for l in range(6):
            worst_g = max(worst_g, np.abs(ctx.gaussian(o, l) - run.gaussian(o, l)).max())
        for l in range(5):
            worst_d = max(worst_d, np.abs(ctx.dog(o, l) - run.dog(o, l)).max())
    REPORT["pyramid_max_abs_err"] = dict(gaussian=worst_g, dog=worst_d)
    assert worst_g < 2e-4 and worst_d < 2e-4
    # golden planes from the real reference
    assert np.abs(ctx.gaussian(1, 3) - synth["g_o1_l3"]).max() < 2e-4
    assert np.abs(ctx.dog(1, 2) - synth["dog_o1_l2"]).max() < 2e-4


def test_stage_lists_match_reference_golden(ctx, synth):
    """Extrema, refined and oriented keypoints vs the real reference's golden dump."""
    img = synth["image"]
    final = ctx.detect(img)
    ex = ctx.extrema()
    both, only_gpu, only_ref = P.set_diff_report(ex, synth["extrema"])
    REPORT["synth_extrema"] = dict(both=both, only_gpu=only_gpu, only_ref=only_ref)
    assert only_gpu + only_ref <= max(2, 0.005 * len(synth["extrema"]))
    for stage, key in ((0, "raw"), (1, "oriented")):
        got = ctx.stage_keypoints(stage)
        rec, prec, gi, wi = P.recall_precision(got, synth[key], use_ori=(stage == 1))
        REPORT[f"synth_{key}"] = dict(n_gpu=len(got), n_ref=len(synth[key]), recall=rec, precision=prec)
        assert rec >= 0.995 and prec >= 0.995
    rec, prec, gi, wi = P.recall_precision(final, synth["final"])
    rep = P.descriptor_parity(final, synth["final"], gi, wi, O.Run(O.best(), img, keep_pyramid=True))
    REPORT["synth_final"] = dict(n_gpu=len(final), n_ref=len(synth["final"]), recall=rec, precision=prec, desc=rep)
    assert rec >= 0.995 and prec >= 0.995
    assert rep["frac_le1"] >= 0.99
    P.assert_descriptor_parity(rep)


@pytest.mark.parametrize("shape_seed", [(192, 256, 42), (300, 400, 31), (480, 640, 77)])
def test_descriptor_stage_on_reference_keypoints(ctx, shape_seed):
    """compute_descriptors (sift.cpp:610-682) + update_histogram (:541-571) + convert_hist_to_desc (:576-603) in
    isolation: the GPU stage on the REFERENCE's own final keypoints (bit-identical x, y, size, pori), over the
    GPU's FP32 scale space, against the reference's descriptors.  north_star bar: max-abs <= 1 level, every
    keypoint, no exceptions."""
    h, w, seed = shape_seed
    img = O.synth_image(h, w, seed=seed)
    run = O.Run(O.best(), img, keep_pyramid=False)
    want = run.keypoints(2)
    ctx.detect(img)
    got = ctx.describe_given(want)
    assert len(want) > 200
    for f in ("x", "y", "size", "pori", "octave", "layer"):
        assert np.array_equal(got[f], want[f]), f
    d = np.abs(got["desc"].astype(np.int16) - want["desc"].astype(np.int16))
    REPORT[f"desc_stage_{w}x{h}"] = dict(n=len(want), max=int(d.max()), frac_exact=float((d.max(1) == 0).mean()),
                                         frac_bins_off_by_one=float((d == 1).mean()))
    assert d.max() <= 1


def test_orientation_stage_on_reference_keypoints(ctx):
    """compute_orientations (sift.cpp:447-533) in isolation: the GPU stage on the reference's raw keypoints.  The
    36-bin histogram rounds every sample's angle to a bin (sift.cpp:489), so a sample within the FP32 scale
    space's error of a bin edge may land next door and shift the interpolated peak slightly: same peaks, pori
    within the pairing tolerance, and bit-identical x / y / size."""
    img = O.synth_image(300, 400, seed=31)
    run = O.Run(O.best(), img, keep_pyramid=False)
    raw, want = run.keypoints(0), run.keypoints(1)
    ctx.detect(img)
    got = ctx.orient_given(raw)
    rec, prec, gi, wi = P.recall_precision(got, want)
    dp = np.abs((got["pori"][gi] - want["pori"][wi] + np.pi) % (2 * np.pi) - np.pi)
    REPORT["orient_stage"] = dict(n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec,
                                  pori_exact=float((dp == 0).mean()), pori_lt_1e6=float((dp < 1e-6).mean()),
                                  pori_max=float(dp.max()))
    assert rec >= 0.995 and prec >= 0.995
    for f in ("x", "y", "size"):
        assert np.array_equal(got[f][gi], want[f][wi]), f
    assert np.quantile(dp, 0.9) < 1e-5


def test_centred_scale_space_is_closer_to_the_fp64_reference(ctx, synth):
    """The scale space is stored relative to the input's mid level (sift_b200_debug_launch_plan): same planes up to
    rounding, with a smaller error against the FP64 oracle than the uncentred FP32 form."""
    img = synth["image"]
    run = O.Run(O.port(), img, keep_pyramid=True)
    err = {}
    for centred in (0, 1):
        ctx.launch_plan(centred=centred)
        ctx.detect(img)
        e = []
        for o in range(3):
            for l in (1, 2, 3):
                e.append(np.abs(ctx.gaussian(o, l).astype(np.float64) - run.gaussian(o, l)).max())
            for l in range(5):
                e.append(np.abs(ctx.dog(o, l).astype(np.float64) - run.dog(o, l)).max())
        err[centred] = float(max(e))
    ctx.launch_plan(centred=1)
    REPORT["centred_vs_plain_max_abs_err"] = err
    assert err[1] <= err[0] and err[1] < 1e-4


@pytest.mark.parametrize("shape", [(192, 256), (301, 517), (97, 1030), (768, 1024), (125, 124), (126, 249)])
def test_extrema_kernel_forms_find_the_same_set(ctx, shape):
    """detect_octave_extrema (sift.cpp:264-291): the four-columns-per-lane kernel (default) and the one-column
    kernel return the same candidate set -- widths around the 124-column strip edge included -- and the same
    final records."""
    h, w = shape
    img = O.synth_image(h, w, seed=h + w)
    ctx.launch_plan(extrema_form=1)
    ref = ctx.detect(img)
    ex1 = ctx.extrema()
    ctx.launch_plan(extrema_form=0)
    got = ctx.detect(img)
    ex4 = ctx.extrema()
    assert len(ex1) > 50
    assert P.set_diff_report(ex4, ex1)[1:] == (0, 0)
    assert len(ex4) == len(ex1)            # no duplicates either
    assert got.tobytes() == ref.tobytes()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_scale_space_kernels_write_exactly_their_planes(ctx, mode):
    """Write audit instead of a sanitizer (compute-sanitizer is closed on the GPU pool): the arena is filled with a
    NaN pattern, then every pyramid kernel form (default mix, per-level, tile cascade, streaming cascade) must
    overwrite every in-image element of the planes it owns and nothing else -- not the row padding, not the planes
    that stay on chip, not the arena behind the last plane.  Sizes with ragged tiles / strips, gray and RGB,
    doubled and not."""
    rng = np.random.default_rng(mode)
    shapes = [(97, 131), (192, 256), (301, 517), (64, 1030), (768, 1024), (33, 41)]
    for h, w in shapes:
        img = O.synth_image(max(h, 8), max(w, 8), seed=h + w)[:h, :w]
        for doubled in (True, False):
            for keep in (False, True):
                ctx.debug_options(keep_all_planes=keep, unfused_pyramid=mode)
                ctx.canary_arm()
                k = ctx.detect(img, double_image_size=doubled)
                stray, missing = ctx.canary_check()
                assert (stray, missing) == (0, 0), (mode, h, w, doubled, keep, stray, missing)
    rgb = rng.integers(0, 256, (150, 210, 3), dtype=np.uint8)
    ctx.debug_options(unfused_pyramid=mode)
    ctx.canary_arm()
    ctx.detect(rgb)
    assert ctx.canary_check() == (0, 0)
    ctx.debug_options()


def test_tail_kernel_equals_octave_by_octave_launches(ctx):
    """k_tail (all small octaves from one ticket counter, hand-over between octaves inside the launch) and the
    multi-octave extrema launch against one launch per octave: every plane and the records bit-identical, repeated
    (the hand-over is a race if it is wrong), with and without the debug planes, graph and plain launches."""
    for h, w, seed in ((300, 400, 5), (97, 131, 1), (768, 1024, 9), (33, 41, 2), (150, 700, 3)):
        img = O.synth_image(max(h, 8), max(w, 8), seed=seed)[:h, :w]
        for keep in (True, False):
            ctx.debug_options(keep_all_planes=keep)
            ctx.tail_kernel(False)
            ref = ctx.detect(img)
            octs = ctx.stats()["octaves"]
            layers = range(6) if keep else range(4)
            planes = [[ctx.gaussian(o, l) for l in layers] + [ctx.dog(o, l) for l in range(5)] for o in range(octs)]
            n_ref = ctx.stats()["extrema"]
            ctx.tail_kernel(True)
            for rep in range(6):
                ctx.launch_plan(use_graph=rep & 1)
                got = ctx.detect(img)
                assert got.tobytes() == ref.tobytes(), (h, w, keep, rep)
                assert ctx.stats()["extrema"] == n_ref
                if rep < 2:
                    for o in range(octs):
                        mine = [ctx.gaussian(o, l) for l in layers] + [ctx.dog(o, l) for l in range(5)]
                        for k, (x, y) in enumerate(zip(mine, planes[o])):
                            assert np.array_equal(x, y), (h, w, keep, o, k)
    ctx.launch_plan(use_graph=1)
    ctx.debug_options()


def test_cube_handover_equals_refinement_loads(synth, monkeypatch):
    """The extrema scan hands the quadratic fit its first 3x3x3 neighbourhood (CandCube); the refinement must get
    the same numbers as when it loads them itself: all off, all on, and only the first 500 candidates handed
    over (the rest take the load path) give the same bytes."""
    imgs = [synth["image"], O.synth_image(301, 517, seed=3), O.synth_image(97, 131, seed=5)]
    out = {}
    for mode in ("0", "1", "500"):
        monkeypatch.setenv("SIFT_B200_CUBES", mode)
        with S.SiftContext(1040, 768) as c:
            out[mode] = [c.detect(im).tobytes() for im in imgs] + [c.detect(imgs[0], double_image_size=False).tobytes()]
    assert out["0"] == out["1"] == out["500"]


def test_graph_replay_equals_plain_launches(ctx, synth):
    """The CUDA-graph launch plan (forked octave chain) and plain single-stream launches run the same kernels on
    the same data: identical bytes; one graph per image size / parameter set, re-captured only on a change."""
    img, img2 = synth["image"], O.synth_image(200, 300, seed=9)
    ctx.launch_plan(use_graph=0)
    ref, ref2 = ctx.detect(img), ctx.detect(img2)
    ctx.launch_plan(use_graph=1)
    g0 = ctx.graphs_built
    a = ctx.detect(img); b = ctx.detect(img); c = ctx.detect(img2); d = ctx.detect(img, peak_ratio=0.7); e = ctx.detect(img)
    assert ctx.graphs_built - g0 == 4          # img, img2, img with other parameters, img again
    assert a.tobytes() == ref.tobytes() and b.tobytes() == ref.tobytes() and e.tobytes() == ref.tobytes()
    assert c.tobytes() == ref2.tobytes()
    assert len(d) != len(a)


def test_output_order_is_the_reference_sort(ctx, synth):
    """clean_keypoints (sift.cpp:20-24): ascending by Keypoint::operator<, no duplicates."""
    k = ctx.detect(synth["image"])
    key = np.stack([k["x"], k["y"], -k["size"], k["pori"], -k["octave"].astype(float)], 1)
    order = np.lexsort(key.T[::-1])
    assert np.array_equal(order, np.arange(len(k)))
    four = np.stack([k["x"], k["y"], k["size"], k["pori"]], 1)
    assert len(np.unique(four, axis=0)) == len(k)


def test_detect_is_bit_reproducible(ctx, synth):
    a = ctx.detect(synth["image"])
    b = ctx.detect(synth["image"])
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("name", ["image1", "image2"])
def test_config1_detect(ctx, config1, golden_dir, name):
    """stitching/image{1,2}.jpg (stb-decoded, RGB): 1286 / 1430 keypoints in the reference."""
    px = np.asarray(Image.open(os.path.join(golden_dir, name + ".png")))
    assert px.ndim == 3 and px.shape[2] == 3
    got = ctx.detect(px)
    want = config1[name + "_final"]
    st = ctx.stats()
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, O.Run(O.best(), px, keep_pyramid=True))
    REPORT[f"config1_{name}"] = dict(stats=st, n_ref=len(want), recall=rec, precision=prec, desc=rep,
                                     ref_extrema=len(config1[name + "_extrema"]), ref_raw=len(config1[name + "_raw"]))
    assert st["octaves"] == 8
    assert abs(st["extrema"] - len(config1[name + "_extrema"])) <= 0.005 * len(config1[name + "_extrema"])
    assert rec >= 0.995 and prec >= 0.995
    assert rep["frac_le1"] >= 0.99
    P.assert_descriptor_parity(rep)


def test_config1_match_on_reference_descriptors(ctx, config1):
    """match_keypoints (sift.cpp:783-815) on the reference's own descriptors: the 269 matches,
    identical indices and distances."""
    a, b = config1["image1_final"]["desc"], config1["image2_final"]["desc"]
    ia, ib, d = ctx.match(a, b)
    assert len(ia) == 269
    assert np.array_equal(ia, config1["match_ia"]) and np.array_equal(ib, config1["match_ib"])
    assert np.array_equal(d, config1["match_dist"])


def test_config1_end_to_end_matches(ctx, golden_dir, config1):
    """detect x2 + match, all on the GPU, vs the reference's match list mapped through the
    keypoint pairing (ratio margins of config 1 are >= 2.5e-3)."""
    k = []
    for name in ("image1", "image2"):
        k.append(ctx.detect(np.asarray(Image.open(os.path.join(golden_dir, name + ".png")))))
    ia, ib, d = ctx.match(k[0]["desc"], k[1]["desc"])
    oa, ob, od = O.match(O.port(), k[0]["desc"], k[1]["desc"])
    assert np.array_equal(ia, oa) and np.array_equal(ib, ob) and np.array_equal(d, od)
    g1, w1 = P.pair_keypoints(k[0], config1["image1_final"])
    g2, w2 = P.pair_keypoints(k[1], config1["image2_final"])
    m1 = dict(zip(g1.tolist(), w1.tolist()))
    m2 = dict(zip(g2.tolist(), w2.tolist()))
    mapped = {(m1.get(int(i), -1), m2.get(int(j), -2)) for i, j in zip(ia, ib)}
    ref = set(zip(config1["match_ia"].tolist(), config1["match_ib"].tolist()))
    REPORT["config1_matches"] = dict(gpu=len(ia), ref=len(ref), common=len(mapped & ref))
    assert len(mapped & ref) >= 0.98 * len(ref)


def test_match_known_answers(ctx, golden_dir):
    """Ties (lowest j wins), duplicate best (never matches), |B| = 0, 1, 2."""
    k = np.load(os.path.join(golden_dir, "match_kat.npz"))
    a, b = k["a"], k["b"]
    cases = {"full": (a, b), "b1": (a[:20], b[:1]), "b0": (a[:20], b[:0]), "a0": (a[:0], b), "b2": (a[:50], b[:2])}
    for tag, (x, y) in cases.items():
        ia, ib, d = ctx.match(x, y)
        assert np.array_equal(ia, k[tag + "_ia"]), tag
        assert np.array_equal(ib, k[tag + "_ib"]), tag
        assert np.array_equal(d, k[tag + "_dist"]), tag


@pytest.mark.parametrize("na,nb", [(1, 1), (33, 1000), (1000, 33), (2500, 3100), (4096, 4096)])
def test_match_random_sizes_vs_oracle(ctx, na, nb):
    a, b = O.synth_descriptors(na, seed=na), O.synth_descriptors(nb, seed=nb + 7)
    b[nb // 2] = b[0]
    if na > 3 and nb > 3:
        b[3] = a[3]
    ia, ib, d = ctx.match(a, b)
    oa, ob, od = O.match(O.port(), a, b)
    assert np.array_equal(ia, oa) and np.array_equal(ib, ob) and np.array_equal(d, od)


def test_match_device_resident_top2(ctx):
    import torch
    a, b = O.synth_descriptors(777, seed=5), O.synth_descriptors(1500, seed=6)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    idx = torch.empty(len(a), dtype=torch.int32, device="cuda")
    d1, d2 = torch.empty_like(idx), torch.empty_like(idx)
    torch.cuda.synchronize()
    ctx.match_enqueue(ta, len(a), tb, len(b), idx, d1, d2)
    ctx.sync()
    D = ((a.astype(np.int64)[:, None, :] - b.astype(np.int64)[None, :, :]) ** 2).sum(-1)
    order = np.argsort(D, axis=1, kind="stable")
    r = np.arange(len(a))
    assert np.array_equal(idx.cpu().numpy(), order[:, 0])
    assert np.array_equal(d1.cpu().numpy(), D[r, order[:, 0]])
    assert np.array_equal(d2.cpu().numpy(), D[r, order[:, 1]])


def test_undoubled_and_f32_and_rgb_inputs(ctx):
    """double_image_size = false (sift.cpp:119-126: same sigma), float input, RGB input."""
    g = O.synth_image(240, 320, seed=3)
    rng = np.random.default_rng(5)
    rgb = np.clip(np.stack([g, np.roll(g, 3, 1), np.roll(g, 5, 0)], -1).astype(np.int32)
                  + rng.integers(-3, 4, (240, 320, 3)), 0, 255).astype(np.uint8)
    for tag, img, doubled in (("gray_undoubled", g, False), ("rgb_doubled", rgb, True),
                              ("rgb_undoubled", rgb, False), ("f32_gray", g.astype(np.float32) * 0.5 + 17.25, True)):
        got = ctx.detect(img, double_image_size=doubled)
        run = O.Run(O.port(), img, doubled, keep_pyramid=True)
        want = run.keypoints(2)
        rec, prec, gi, wi = P.recall_precision(got, want)
        rep = P.descriptor_parity(got, want, gi, wi, run)
        REPORT[tag] = dict(n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep)
        assert len(want) > 50
        assert rec >= 0.99 and prec >= 0.99, tag
        P.assert_descriptor_parity(rep)


def test_max_octaves_extension_filters_octaves(ctx, synth):
    """Octave o never depends on octaves > o (sift.cpp:187-199): capping = filtering."""
    full = ctx.detect(synth["image"])
    capped = ctx.detect(synth["image"], max_octaves=4)
    want = full[full["octave"] < 4]
    assert capped.tobytes() == want.tobytes()
    assert ctx.stats()["octaves"] == 4


def test_small_and_odd_sizes(ctx):
    for h, w in ((2, 2), (3, 5), (7, 9), (17, 33), (64, 31), (129, 257)):
        img = O.synth_image(max(h, 8), max(w, 8), seed=h * w)[:h, :w]
        got = ctx.detect(img)
        run = O.Run(O.port(), img, keep_pyramid=True)
        want = run.keypoints(2)
        assert ctx.stats()["octaves"] == run.octaves, (h, w)
        rec, prec, _, _ = P.recall_precision(got, want)
        assert len(want) - rec * len(want) <= 1 and len(got) - prec * len(got) <= 1, (h, w)


def test_flat_image_has_no_keypoints(ctx):
    assert len(ctx.detect(np.full((100, 120), 128, np.uint8))) == 0
    assert ctx.stats()["extrema"] == 0


def test_error_behaviour(ctx):
    img = O.synth_image(64, 64, seed=1)
    for bad in (dict(intervals=1), dict(intervals=6), dict(window_size=4), dict(window_size=9), dict(num_bins=2),
                dict(window_size=5, intervals=2),        # no DoG layer left between the borders
                dict(intervals=2, init_sigma=3.0)):      # blur radius beyond the instantiated kernels
        with pytest.raises(S.SiftError) as e:
            ctx.detect(img, **bad)
        assert e.value.code == 5, bad
    with pytest.raises(S.SiftError) as e:
        ctx.detect(np.zeros((64, 64, 2), np.uint8))
    assert e.value.code == 1
    with pytest.raises(S.SiftError) as e:
        ctx.detect(np.zeros((4000, 4000), np.uint8))
    assert e.value.code == 6
    with pytest.raises(S.SiftError) as e:
        ctx.detect(O.synth_image(192, 256, seed=42), capacity=10)
    assert e.value.code == 4


def test_device_resident_enqueue_path(ctx, synth):
    import torch
    img = torch.from_numpy(synth["image"]).cuda()
    torch.cuda.synchronize()
    ctx.detect_enqueue(img, 256, 192)
    n = ctx.detect_finish()
    host = ctx.detect(synth["image"])
    assert n == len(host)
    rec, desc, n2 = ctx.result_device()
    assert n2 == n and rec and desc


@pytest.fixture
def force_path():
    def _set(p):
        if p is None:
            os.environ.pop("SIFT_B200_MATCH", None)
        else:
            os.environ["SIFT_B200_MATCH"] = p
    yield _set
    os.environ.pop("SIFT_B200_MATCH", None)


def _numpy_top2(a, b):
    D = ((a.astype(np.int64)[:, None, :] - b.astype(np.int64)[None, :, :]) ** 2).sum(-1)
    order = np.argsort(D, axis=1, kind="stable")
    r = np.arange(len(a))
    second = D[r, order[:, 1]] if b.shape[0] > 1 else np.full(len(a), 2 ** 31 - 1)
    return order[:, 0], D[r, order[:, 0]], second


@pytest.mark.parametrize("na,nb", [(1, 1), (5, 2), (128, 256), (129, 257), (300, 255), (1000, 1500), (2500, 3100)])
@pytest.mark.parametrize("path", ["tc", "simt"])
def test_match_paths_top2_exact(ctx, force_path, path, na, nb):
    """Both matcher kernels (tcgen05 kind::i8 and SIMT dp4a) return the exact integer
    (nearest index, d1^2, d2^2) of sift.cpp:789-806, including ties (lowest j) and duplicates."""
    import torch
    force_path(path)
    assert ctx.match_path(na, nb) == (1 if path == "tc" else 0)
    a, b = O.synth_descriptors(na, seed=na + 1), O.synth_descriptors(nb, seed=nb + 2)
    if nb > 40:
        b[nb - 1] = b[3]            # duplicate rows far apart (different tiles / splits)
        b[nb // 2] = a[na // 2]     # exact hit
        b[7] = a[0]; b[nb - 2] = a[0]   # two exact hits: best = lowest j, second distance 0
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    idx = torch.full((na,), -7, dtype=torch.int32, device="cuda")
    d1, d2 = torch.zeros_like(idx), torch.zeros_like(idx)
    torch.cuda.synchronize()
    ctx.match_enqueue(ta, na, tb, nb, idx, d1, d2)
    ctx.sync()
    wi, w1, w2 = _numpy_top2(a, b)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(d1.cpu().numpy(), w1)
    assert np.array_equal(d2.cpu().numpy(), w2)


def test_match_tc_extreme_values(ctx, force_path):
    """All-255 against all-0 rows: the largest possible distance (128 * 255^2) and dot product."""
    import torch
    force_path("tc")
    a = np.zeros((256, 128), np.uint8); a[::2] = 255
    b = np.zeros((512, 128), np.uint8); b[1::2] = 255; b[5, :64] = 255
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    idx = torch.empty(256, dtype=torch.int32, device="cuda")
    d1, d2 = torch.empty_like(idx), torch.empty_like(idx)
    torch.cuda.synchronize()
    ctx.match_enqueue(ta, 256, tb, 512, idx, d1, d2)
    ctx.sync()
    wi, w1, w2 = _numpy_top2(a, b)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(d1.cpu().numpy(), w1) and np.array_equal(d2.cpu().numpy(), w2)


def test_match_large_tc_vs_simt(ctx, force_path):
    """20k x 20k (one pair of config 5): the two kernels agree bit for bit."""
    import torch
    a, b = O.synth_descriptors(20000, seed=11), O.synth_descriptors(20000, seed=12)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = {}
    for path in ("tc", "simt"):
        force_path(path)
        idx = torch.empty(20000, dtype=torch.int32, device="cuda")
        d1, d2 = torch.empty_like(idx), torch.empty_like(idx)
        torch.cuda.synchronize()
        ctx.match_enqueue(ta, 20000, tb, 20000, idx, d1, d2)
        ctx.sync()
        out[path] = (idx.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy())
    for x, y in zip(out["tc"], out["simt"]):
        assert np.array_equal(x, y)


def test_fused_octave_cascade_equals_per_level_kernels(ctx):
    """The fused per-octave cascades -- tile kernels (k_cascade, mode 2), streaming kernels (k_stream,
    mode 3) and the default mix of the two (mode 0) -- and the one-kernel-per-level path (mode 1) run
    the same arithmetic in the same order: every plane and the final records are bit-identical,
    including image sizes that are not multiples of the tile / strip and octaves smaller than one."""
    for h, w, seed in ((192, 256, 42), (301, 517, 7), (97, 1030, 8), (768, 1024, 9)):
        img = O.synth_image(h, w, seed=seed)
        ctx.debug_options(keep_all_planes=True, unfused_pyramid=1)
        ref = ctx.detect(img)
        octs = ctx.stats()["octaves"]
        planes = [[ctx.gaussian(o, l) for l in range(6)] + [ctx.dog(o, l) for l in range(5)] for o in range(octs)]
        for mode in (0, 2, 3):
            ctx.debug_options(keep_all_planes=True, unfused_pyramid=mode)
            got = ctx.detect(img)
            for o in range(octs):
                mine = [ctx.gaussian(o, l) for l in range(6)] + [ctx.dog(o, l) for l in range(5)]
                for k, (x, y) in enumerate(zip(mine, planes[o])):
                    assert np.array_equal(x, y), (mode, h, w, o, k, float(np.abs(x - y).max()))
            assert got.tobytes() == ref.tobytes(), (mode, h, w)
    ctx.debug_options()


def test_profile_marks_add_up_to_the_stage_profile(ctx):
    """sift_b200_get_profile_marks is the unaggregated form of sift_b200_get_profile: same stages, same total
    time, the fused pyramid bracketed kernel by kernel on octave 0 and per octave below."""
    img = O.synth_image(300, 400, seed=5)
    ctx.set_profiling(True)
    ctx.tail_kernel(False)
    ctx.detect(img)
    ms, nl = ctx.profile()
    marks = ctx.profile_marks()
    octs = ctx.stats()["octaves"]
    assert [st for st, _ in marks][:4] == ["input", "pyramid", "pyramid", "extrema"]
    assert sum(1 for st, _ in marks if st == "pyramid") == octs + 1
    assert nl["pyramid"] == 2 * octs
    ctx.tail_kernel(True)      # octaves 1.. of the 600 x 800 base are at most one tile per SM: one launch for all
    ctx.detect(img)
    ms, nl = ctx.profile()
    marks = ctx.profile_marks()
    ctx.set_profiling(False)
    assert nl["pyramid"] == 3 and nl["extrema"] == 2
    assert [st for st, _ in marks][:6] == ["input", "pyramid", "pyramid", "extrema", "pyramid", "extrema"]
    for st in S.SiftContext.STAGES:
        assert abs(sum(t for s_, t in marks if s_ == st) - ms[st]) < 1e-3, st
    assert all(t >= 0 for _, t in marks)


def test_streaming_cascade_on_a_large_octave():
    """Default mode at a size where octave 0 takes the streaming kernels (>= 2 Mpx) and the rest the tile
    kernels: same bytes as tile kernels everywhere and as streaming kernels everywhere, without the debug
    planes (G4, G5 stay on chip).  Width 3000 -> base 6000 x 3600: strips of 96 / 232 columns do not divide it."""
    img = O.synth_image(1800, 3000, seed=11)
    with S.SiftContext(3000, 1800) as c:
        c.debug_options(unfused_pyramid=2)
        ref = c.detect(img)
        c.debug_options(unfused_pyramid=0)
        got = c.detect(img)
        c.debug_options(unfused_pyramid=3)
        got3 = c.detect(img)
    assert len(ref) > 5000
    assert got.tobytes() == ref.tobytes()
    assert got3.tobytes() == ref.tobytes()


# ------------------------------------------------------------------------------------------
# BASELINE.json configs at full size.  Known answers of the REAL reference (SURVEY.md section 4,
# measured with oracle/_ref on the same generator-D images): they pin the GPU path at sizes the CPU
# oracle needs minutes for.
# ------------------------------------------------------------------------------------------
KNOWN = {  # (height, width, seed) -> (octaves, extrema, raw, oriented, final)
    (1080, 1920, 1234): (9, 39448, 5776, 7259, 7240),
    (2160, 3840, 1234): (10, 157361, 21887, 27555, 27536),
}


@pytest.mark.parametrize("shape", sorted(KNOWN))
def test_full_size_stage_counts_match_reference_known_answers(shape):
    h, w, seed = shape
    octs, n_ext, n_raw, n_ori, n_fin = KNOWN[shape]
    img = O.synth_image(h, w, seed=seed)
    with S.SiftContext(w, h) as c:
        k = c.detect(img)
        st = c.stats()
    REPORT[f"known_{w}x{h}"] = dict(stats=st, reference=dict(extrema=n_ext, raw=n_raw, oriented=n_ori, final=n_fin))
    assert st["octaves"] == octs
    for got, want in ((st["extrema"], n_ext), (st["raw_keypoints"], n_raw), (st["oriented_keypoints"], n_ori),
                      (len(k), n_fin)):
        assert abs(got - want) <= 0.005 * want, (got, want)   # 99.5 % set agreement implies this
    # the reference's output order and de-duplication
    key = np.stack([k["x"], k["y"], -k["size"], k["pori"], -k["octave"].astype(float)], 1)
    assert np.array_equal(np.lexsort(key.T[::-1]), np.arange(len(k)))


def test_config2_1080p_four_octaves_vs_oracle():
    """Config 2: 1920x1080, 4 octaves x 5 scales.  The reference derives the octave count, so the
    oracle is its result filtered to octave < 4 (octave o never depends on octaves > o)."""
    img = O.synth_image(1080, 1920, seed=1234)
    with S.SiftContext(1920, 1080) as c:
        got = c.detect(img, max_octaves=4)
    run = O.Run(O.best(), img, keep_pyramid=True)
    want = run.keypoints(2)
    want = want[want["octave"] < 4]
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    REPORT["config2_1080p_4oct"] = dict(n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep)
    assert rec >= 0.995 and prec >= 0.995
    assert rep["frac_le1"] >= 0.99
    P.assert_descriptor_parity(rep)


def _large_golden(golden_dir, w, h, seed=1234):
    """Final keypoints of the REAL reference at full size (tests/golden/make_golden_large.py)."""
    z = np.load(os.path.join(golden_dir, f"synth_{w}x{h}_seed{seed}.npz"))
    k = np.zeros(len(z["x"]), dtype=S.KP_DTYPE)
    for f in ("x", "y", "size", "pori", "octave", "layer"):
        k[f] = z[f]
    stride = int(z["desc_stride"])
    k["desc"][::stride] = z["desc"]
    return k, z["counts"], stride, int(z["image_crc"])


def _full_size_set_parity(tag, got, want, stride, ctx):
    """Set-level comparison with the reference's keypoints (north_star: recall / precision >= 99.5 % at 0.01 px
    and 1e-3 relative size), descriptors of the paired keypoints, and the descriptor stage alone on the
    reference's own keypoints (max-abs <= 1 level)."""
    rec, prec, gi, wi = P.recall_precision(got, want)
    has = (wi % stride) == 0                       # pairs whose reference descriptor is in the fixture
    rep = P.descriptor_report(got, want, gi[has], wi[has])
    d = np.abs(got["desc"][gi[has]].astype(np.int16) - want["desc"][wi[has]].astype(np.int16)).max(1)
    g, w_ = got[gi[has]][d > 1], want[wi[has]][d > 1]
    same = (g["x"] == w_["x"]) & (g["y"] == w_["y"]) & (g["size"] == w_["size"]) & (g["pori"] == w_["pori"])
    stage = ctx.describe_given(want[::stride])     # GPU descriptor stage on the reference's keypoints
    sd = np.abs(stage["desc"].astype(np.int16) - want["desc"][::stride].astype(np.int16))
    REPORT[tag] = dict(n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep,
                       n_outliers=int((d > 1).sum()), unexplained=int(same.sum()),
                       desc_stage=dict(n=len(stage), max=int(sd.max()), frac_exact=float((sd.max(1) == 0).mean())))
    assert rec >= 0.995 and prec >= 0.995, (rec, prec)
    assert rep["frac_le1"] >= 0.99, rep
    assert same.sum() == 0                          # a differing descriptor always comes with a differing keypoint
    assert sd.max() <= 1


def test_config3_4k_set_parity_vs_reference(golden_dir):
    """Config 3's image 0 (3840x2160, generator D, seed 1234) against EVERY final keypoint of the real reference
    (detect_keypoints_and_descriptors, sift.cpp:712-776: 27 536 keypoints, ~104 s of CPU, run once by
    tests/golden/make_golden_large.py)."""
    want, counts, stride, crc = _large_golden(golden_dir, 3840, 2160)
    img = O.synth_image(2160, 3840, seed=1234)
    assert int(img.astype(np.uint64).sum()) == crc      # same pixels as the fixture's run
    with S.SiftContext(3840, 2160) as c:
        got = c.detect(img)
        st = c.stats()
        assert [st["octaves"], len(got)] == [10, st["final_keypoints"]]
        _full_size_set_parity("config3_4k_set", got, want, stride, c)
    assert list(counts[1:]) == [157361, 21887, 27555, 27536]   # extrema / raw / oriented / final, SURVEY.md section 4


def test_config4_8k_set_parity_vs_reference(golden_dir):
    """Config 4 (7680x4320, 11 octaves, base 15360x8640): every final keypoint of the real reference; descriptors of
    every 8th (fixture size)."""
    want, counts, stride, crc = _large_golden(golden_dir, 7680, 4320)
    img = O.synth_image(4320, 7680, seed=1234)
    assert int(img.astype(np.uint64).sum()) == crc
    with S.SiftContext(7680, 4320) as c:
        got = c.detect(img)
        st = c.stats()
        assert st["octaves"] == 11 and st["base_width"] == 15360
        assert abs(st["extrema"] - counts[1]) <= 0.005 * counts[1]
        _full_size_set_parity("config4_8k_set", got, want, stride, c)


def test_config4_8k_properties():
    """Config 4 (7680x4320) on the benchmark's GPU generator: size-independent properties."""
    import torch
    sys_path_bench = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(sys_path_bench, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    img = bench.synth_image_gpu(4320, 7680, 1234, torch.device("cuda", 0))
    torch.cuda.synchronize()
    with S.SiftContext(7680, 4320) as c:
        c.detect_enqueue(img, 7680, 4320)
        n = c.detect_finish()
        out = np.zeros(n, dtype=S.KP_DTYPE)
        assert c.result_copy(out) == n
        st = c.stats()
        c.detect_enqueue(img, 7680, 4320)
        out2 = np.zeros(n, dtype=S.KP_DTYPE)
        assert c.detect_finish() == n and c.result_copy(out2) == n
    REPORT["config4_8k"] = dict(stats=st)
    assert st["octaves"] == 11 and st["base_width"] == 15360
    assert 60000 < n < 200000                      # ~4 k keypoints per input megapixel
    assert out.tobytes() == out2.tobytes()          # bit-reproducible
    key = np.stack([out["x"], out["y"], -out["size"], out["pori"], -out["octave"].astype(float)], 1)
    assert np.array_equal(np.lexsort(key.T[::-1]), np.arange(n))
    assert out["x"].min() >= 0 and out["x"].max() < 7680 and out["y"].max() < 4320
    assert set(np.unique(out["layer"])) <= {1, 2, 3}
    assert (out["desc"].astype(np.float64) ** 2).sum(1).min() > 0   # no empty descriptors


def test_collection_through_the_c_abi_on_one_gpu(ctx):
    """sift_b200_collection_match with world = 1 (no NCCL): every unordered pair once, in both directions on
    request, empty sets included; each pair's matches equal match_keypoints (sift.cpp:783-815) on that pair."""
    sets = [O.synth_descriptors(n, seed=200 + k) for k, n in enumerate((700, 0, 1, 333, 1500, 2))]
    sets[4][10] = sets[0][5]
    for both in (False, True):
        n_pairs = ctx.collection_match(len(sets), sets, both_directions=both)
        pi, pj, rows = ctx.collection_pairs()
        want = S.partition_pairs_native([len(s) for s in sets], 1, 0, both)
        assert n_pairs == len(want) == (30 if both else 15)
        assert sorted(zip(pi.tolist(), pj.tolist())) == sorted(want)
        total = 0
        for q, (i, j, r) in enumerate(zip(pi, pj, rows)):
            assert r == len(sets[i])
            ia, ib, d = ctx.collection_fetch(q, r)
            wa, wb, wd = O.match(O.port(), sets[i], sets[j])
            assert np.array_equal(ia, wa) and np.array_equal(ib, wb) and np.array_equal(d, wd), (i, j)
            total += len(ia)
        m, h = ctx.collection_digest(all_ranks=True)
        assert m == total
        assert ctx.collection_digest(all_ranks=False) == (m, h)      # deterministic


def test_batch_detect_over_several_contexts(synth):
    """sift_b200_detect_batch_u8: image k on context k % n; same bytes as one call per image."""
    imgs = [synth["image"], O.synth_image(192, 256, seed=1), O.synth_image(192, 256, seed=2),
            O.synth_image(192, 256, seed=3), O.synth_image(192, 256, seed=4)]
    ctxs = [S.SiftContext(256, 192) for _ in range(3)]
    got = S.detect_batch(ctxs, imgs)
    for im, g in zip(imgs, got):
        assert g.tobytes() == ctxs[0].detect(im).tobytes()
    with pytest.raises(S.SiftError) as e:
        S.detect_batch(ctxs, imgs, capacity=10)
    assert e.value.code == 4
    for c in ctxs:
        c.close()


def test_contexts_come_and_go_without_leaking_and_sizes_may_alternate():
    """Housekeeping of the host layer: contexts created and destroyed in a loop give the device memory back (arena,
    lists, graph, streams, events); one context fed alternating image sizes and parameters re-captures its graph and
    keeps returning the bytes a fresh context returns; several contexts interleaved on one GPU do not disturb each
    other."""
    import torch
    imgs = [O.synth_image(120, 160, seed=1), O.synth_image(200, 150, seed=2), O.synth_image(96, 300, seed=3)]
    with S.SiftContext(320, 240) as c:
        want = [c.detect(im) for im in imgs]
        want_p = c.detect(imgs[0], peak_ratio=0.7)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for rep in range(12):
        with S.SiftContext(320, 240) as c:
            for k in (0, 1, 0, 2, 1):
                assert c.detect(imgs[k]).tobytes() == want[k].tobytes(), (rep, k)
            assert c.detect(imgs[0], peak_ratio=0.7).tobytes() == want_p.tobytes()
            assert c.detect(imgs[0]).tobytes() == want[0].tobytes()
            a, b = O.synth_descriptors(400, seed=rep), O.synth_descriptors(900, seed=rep + 50)
            c.match(a, b)
            c.collection_match(2, [a, b])
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20, (free0, free1)      # nothing accumulates (allocator granularity aside)
    ctxs = [S.SiftContext(320, 240) for _ in range(4)]
    for rep in range(3):
        for k, c in enumerate(ctxs):
            c.detect_enqueue(np.ascontiguousarray(imgs[(k + rep) % 3]), imgs[(k + rep) % 3].shape[1], imgs[(k + rep) % 3].shape[0])
        for k, c in enumerate(ctxs):
            n = c.detect_finish()
            out = np.zeros(n, dtype=S.KP_DTYPE)
            assert c.result_copy(out) == n
            assert out.tobytes() == want[(k + rep) % 3].tobytes()
    for c in ctxs:
        c.close()


def test_match_self_is_identity(ctx):
    """match(A, A): every row's nearest neighbour is itself at distance 0 (any size, both kernels)."""
    import torch
    a = O.synth_descriptors(5000, seed=77)
    ta = torch.from_numpy(a).cuda()
    idx = torch.empty(5000, dtype=torch.int32, device="cuda")
    d1, d2 = torch.empty_like(idx), torch.empty_like(idx)
    torch.cuda.synchronize()
    ctx.match_enqueue(ta, 5000, ta, 5000, idx, d1, d2)
    ctx.sync()
    assert np.array_equal(idx.cpu().numpy(), np.arange(5000))
    assert int(d1.max()) == 0 and int(d2.min()) > 0


@pytest.mark.parametrize("kw", [
    dict(init_sigma=1.3, contrast_threshold=0.03, eigen_ratio=6.0, peak_ratio=0.7, ori_sigma_factor=1.2,
         desc_scale_factor=2.5),
    dict(double_image_size=False, init_sigma=2.0, contrast_threshold=0.08, eigen_ratio=15.0, peak_ratio=0.9,
         ori_sigma_factor=1.8, desc_scale_factor=3.5),
    dict(contrast_threshold=0.02, eigen_ratio=20.0, peak_ratio=0.6),
])
def test_non_default_arguments_vs_oracle(ctx, kw):
    """The tunable arguments of sift.hh:65-71 (other sigmas take the per-level blur kernels instead
    of the fused cascade)."""
    img = O.synth_image(300, 400, seed=31)
    got = ctx.detect(img, **kw)
    run = O.Run(O.best(), img, params=O.Params(**kw), keep_pyramid=True)
    want = run.keypoints(2)
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    REPORT["params_" + "_".join(f"{k}={v}" for k, v in sorted(kw.items()))[:60]] = dict(
        n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep)
    assert len(want) > 30
    assert rec >= 0.99 and prec >= 0.99
    assert rep["frac_le1"] >= 0.98
    P.assert_descriptor_parity(rep)


@pytest.mark.parametrize("kw", [dict(intervals=2, contrast_threshold=0.05), dict(intervals=4, init_sigma=1.4, peak_ratio=0.75),
                                dict(intervals=5), dict(intervals=2, double_image_size=False)])
def test_other_interval_counts_vs_oracle(ctx, kw):
    """intervals != 3: 5..8 Gaussian and 4..7 DoG layers per octave, the threshold
    floor(0.5 ct / intervals * 255), the size formula and the next-octave base G[intervals] all follow."""
    img = O.synth_image(300, 400, seed=33)
    got = ctx.detect(img, **kw)
    run = O.Run(O.best(), img, params=O.Params(**kw), keep_pyramid=True)
    want = run.keypoints(2)
    st = ctx.stats()
    assert st["octaves"] == run.octaves
    both, only_gpu, only_ref = P.set_diff_report(ctx.extrema(), run.extrema().astype(np.int64))
    assert only_gpu + only_ref <= max(2, 0.005 * (both + only_ref))
    L = kw["intervals"] + 3
    worst = max(np.abs(ctx.dog(1, l) - run.dog(1, l)).max() for l in range(L - 1))
    assert worst < 3e-4
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    REPORT["intervals_" + "_".join(f"{k}={v}" for k, v in sorted(kw.items()))[:50]] = dict(
        n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep, dog_err=worst)
    assert len(want) > 30
    assert rec >= 0.99 and prec >= 0.99
    assert rep["frac_le1"] >= 0.98
    P.assert_descriptor_parity(rep)


def test_random_shapes_fused_equals_per_level(ctx):
    """A sweep of awkward shapes (primes, thin strips, sizes around the 64 / 128 tile edges): the
    fused input + cascade kernels and the per-level kernels must give identical bytes."""
    rng = np.random.default_rng(2024)
    shapes = [(67, 131), (64, 64), (65, 129), (127, 63), (129, 65), (40, 700), (500, 37), (193, 257), (255, 511)]
    shapes += [(int(rng.integers(33, 400)), int(rng.integers(33, 640))) for _ in range(6)]
    for h, w in shapes:
        img = O.synth_image(h, w, seed=h * 1000 + w)
        for doubled in (True, False):
            ctx.debug_options(unfused_pyramid=1)
            ref = ctx.detect(img, double_image_size=doubled)
            for mode in (2, 3):   # tile cascade, streaming cascade
                ctx.debug_options(unfused_pyramid=mode)
                got = ctx.detect(img, double_image_size=doubled)
                assert got.tobytes() == ref.tobytes(), (mode, h, w, doubled)
    ctx.debug_options()


def test_rgb_input_fused_equals_unfused(ctx):
    """convert_to_grayscale + resize_inter_bilinear + the initial blur (image.cpp:8-24, 62-88, sift.cpp:113-126) for
    RGB input: the fused input kernel (gray in FP64 exactly as the reference forms it, tile formed in shared memory)
    and the unfused k_prepare + k_blur path give the same bytes, doubled or not, edges included."""
    rng = np.random.default_rng(12)
    for h, w in ((97, 131), (192, 256), (300, 413)):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        g = O.synth_image(h, w, seed=h)
        rgb = np.clip(rgb.astype(np.int32) // 8 + g[..., None].astype(np.int32) - 16, 0, 255).astype(np.uint8)
        for doubled in (True, False):
            ctx.debug_options(unfused_pyramid=1)
            ref = ctx.detect(rgb, double_image_size=doubled)
            base_ref = ctx.gaussian(0, 0)
            ctx.debug_options(unfused_pyramid=0)
            got = ctx.detect(rgb, double_image_size=doubled)
            assert np.array_equal(ctx.gaussian(0, 0), base_ref), (h, w, doubled)
            assert got.tobytes() == ref.tobytes() and len(got) > 20, (h, w, doubled)
    ctx.debug_options()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_small_images_vs_oracle(ctx, seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(90, 260)), int(rng.integers(90, 330))
    img = O.synth_image(h, w, seed=seed + 500)
    got = ctx.detect(img)
    run = O.Run(O.best(), img, keep_pyramid=True)
    want = run.keypoints(2)
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    assert rec >= 0.99 and prec >= 0.99, (h, w, rec, prec)
    assert rep["frac_le1"] >= 0.97
    P.assert_descriptor_parity(rep)


def test_noise_and_dense_texture_images(ctx):
    """iid noise (sparse: the DoG of white noise rarely clears the contrast threshold) and a fine
    texture with ~4x the keypoint density of generator D (SURVEY.md 8d calibration)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (384, 512), dtype=np.uint8)
    acc = np.zeros((384, 512))
    for s in (1, 2, 4, 8, 16):
        n = gaussian_filter(rng.standard_normal((384, 512)), s, mode="wrap")
        acc += n / n.std()
    dense = np.rint((acc - acc.min()) / (acc.max() - acc.min()) * 255.0).astype(np.uint8)
    for tag, img in (("iid_noise", noise), ("dense_texture", dense)):
        got = ctx.detect(img)
        st = ctx.stats()
        want = O.Run(O.best(), img, keep_pyramid=False).keypoints(2)
        rec, prec, gi, wi = P.recall_precision(got, want)
        REPORT[tag] = dict(stats=st, n_ref=len(want), recall=rec, precision=prec)
        assert rec >= 0.99 and prec >= 0.99, tag
    assert st["final_keypoints"] > 2000


@pytest.mark.parametrize("kw", [dict(num_bins=18), dict(num_bins=72, peak_ratio=0.7), dict(window_size=5, intervals=4),
                                dict(window_size=5, contrast_threshold=0.02)])
def test_window_and_bin_counts_vs_oracle(ctx, kw):
    """window_size and num_bins: (2b+1)^3 tie-tolerant extrema with border b in the
    scan and in the refinement bounds; orientation histograms with other bin counts."""
    img = O.synth_image(300, 400, seed=35)
    got = ctx.detect(img, **kw)
    run = O.Run(O.best(), img, params=O.Params(**kw), keep_pyramid=True)
    want = run.keypoints(2)
    both, only_gpu, only_ref = P.set_diff_report(ctx.extrema(), run.extrema().astype(np.int64))
    assert only_gpu + only_ref <= max(2, 0.005 * (both + only_ref))
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    REPORT["knobs_" + "_".join(f"{k}={v}" for k, v in sorted(kw.items()))[:50]] = dict(
        n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep)
    assert len(want) > 30
    assert rec >= 0.99 and prec >= 0.99
    assert rep["frac_le1"] >= 0.98
    P.assert_descriptor_parity(rep)


@pytest.mark.parametrize("scale,offset", [(100.0, 0.0), (0.01, -1.0), (1.0, -128.0)])
def test_float_input_of_any_range(ctx, scale, offset):
    """sift_b200_detect_f32 on data far outside 0..255 (and negative): the fixed-point histogram
    scale follows the input's own range, so nothing overflows or underflows."""
    img = O.synth_image(240, 320, seed=41).astype(np.float32) * np.float32(scale) + np.float32(offset)
    kw = dict(contrast_threshold=0.04 * scale) if scale < 1 else {}
    got = ctx.detect(img, **kw)
    run = O.Run(O.best(), img.astype(np.float64), params=O.Params(**kw), keep_pyramid=True)
    want = run.keypoints(2)
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_parity(got, want, gi, wi, run)
    REPORT[f"f32_range_x{scale}_{offset}"] = dict(n_gpu=len(got), n_ref=len(want), recall=rec, precision=prec, desc=rep)
    assert len(want) > 50
    assert rec >= 0.99 and prec >= 0.99
    assert rep["frac_le1"] >= 0.97
    P.assert_descriptor_parity(rep)
