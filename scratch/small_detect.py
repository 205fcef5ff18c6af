"""Smallest end-to-end exercise of every kernel family (target for compute-sanitizer): one small detect in each
pyramid mode (per-level, tile cascade, streaming cascade, default), RGB + float inputs, both extrema forms, the
graph plan, a tensor-core and a SIMT match, a 1-GPU collection match."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sift_project_b200 as S
from oracle import oracle as O

img = O.synth_image(96, 160, seed=3)
rgb = np.stack([img, np.roll(img, 2, 1), np.roll(img, 3, 0)], -1)
with S.SiftContext(160, 96) as c:
    ref = None
    for graph in (0, 1):
        c.launch_plan(use_graph=graph)
        for mode in (1, 2, 3, 0):
            c.debug_options(unfused_pyramid=mode)
            k = c.detect(img)
            ref = k if ref is None else ref
            assert k.tobytes() == ref.tobytes(), (graph, mode)
    c.debug_options()
    c.launch_plan(extrema_form=1); assert c.detect(img).tobytes() == ref.tobytes()
    c.launch_plan(extrema_form=0)
    n_rgb = len(c.detect(rgb)); n_f32 = len(c.detect(img.astype(np.float32) * 0.5 + 3.0)); n_und = len(c.detect(img, double_image_size=False))
    a, b = O.synth_descriptors(300, seed=1), O.synth_descriptors(700, seed=2)
    os.environ["SIFT_B200_MATCH"] = "tc"; m1 = c.match(a, b)
    os.environ["SIFT_B200_MATCH"] = "simt"; m2 = c.match(a, b)
    os.environ.pop("SIFT_B200_MATCH")
    assert all(np.array_equal(x, y) for x, y in zip(m1, m2))
    n_pairs = c.collection_match(3, [a, b, a[:50]])
    dig = c.collection_digest(all_ranks=False)
    st = c.describe_given(ref)
    print("ok:", len(ref), "keypoints;", n_rgb, n_f32, n_und, "rgb / f32 / undoubled;", len(m1[0]), "matches;", n_pairs, "pairs, digest", dig)
