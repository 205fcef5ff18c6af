"""Collection (all-pairs) matching across the GPUs of one box -- BASELINE.json config 5.

The reference matches two images (src/main.cpp:14-17 -> match_keypoints, sift.cpp:783-815); a
stitching collection repeats that for every image pair.  Here every rank owns the descriptors of
its images (detect is sharded by image, no collective); ONE exchange step -- an all-gather of the
u8 descriptor blocks (NCCL over NVLink on GPUs, gloo in the CPU tests) -- gives every rank the whole
collection (512 x 20k x 128 B = 1.31 GB for config 5), after which the image pairs are dealt to
ranks by flop weight and matched locally; a pair is never split across ranks, so no reduction
collective is needed.  Only the (small) match lists travel to rank 0.

torch.distributed is plumbing; the matching itself is libsift_b200's tcgen05 kernel
(SiftContext.match).  `matcher` exists so that the CPU tests can drive the host logic with the
oracle as a stand-in -- the product default has no CPU path.
"""
import numpy as np


def pair_list(n_images, both_directions=False):
    pairs = [(i, j) for i in range(n_images) for j in range(i + 1, n_images)]
    if both_directions:
        pairs += [(j, i) for (i, j) in pairs]
    return pairs


def partition_pairs(counts, world, both_directions=False):
    """Deterministic longest-processing-time assignment of image pairs to ranks.

    cost(i, j) = counts[i] * counts[j] (the matcher's 2 * n_i * n_j * 128 flop); pairs are taken
    in descending cost (ties: lexicographic) and given to the least-loaded rank (ties: lowest
    rank).  Returns a list of pair lists, one per rank; every pair appears exactly once."""
    counts = [int(c) for c in counts]
    pairs = pair_list(len(counts), both_directions)
    pairs.sort(key=lambda p: (-counts[p[0]] * counts[p[1]], p))
    load = [0] * world
    out = [[] for _ in range(world)]
    for p in pairs:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(p)
        load[r] += counts[p[0]] * counts[p[1]]
    return out


def owner_of(image, world):
    """Detect is sharded round-robin by image (bench.py)."""
    return image % world


def all_gather_descriptors(local, n_images, group=None, device=None):
    """local: {image index: (n_i, 128) uint8 torch tensor} for the images this rank owns.
    Returns (list of n_images tensors on this rank, counts).  One all-gather of counts (tiny) and
    one of the padded descriptor blocks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = sorted(local)
    assert all(owner_of(i, world) == rank for i in mine), "a rank may only hold the images it owns"
    if device is None:
        if local:
            device = next(iter(local.values())).device
        elif dist.is_initialized() and dist.get_backend(group) == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())   # a rank may own no image
        else:
            device = torch.device("cpu")
    counts = torch.zeros(n_images, dtype=torch.int64, device=device)
    for i in mine:
        counts[i] = local[i].shape[0]
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts_h = counts.cpu().tolist()
    per_rank = [sum(counts_h[i] for i in range(n_images) if owner_of(i, world) == r) for r in range(world)]
    rows = max(max(per_rank), 1)
    block = torch.zeros((rows, 128), dtype=torch.uint8, device=device)
    off = 0
    for i in mine:
        n = local[i].shape[0]
        block[off:off + n] = local[i].to(device)
        off += n
    if world > 1:
        gathered = torch.empty((world, rows, 128), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered.view(world * rows, 128), block, group=group)
    else:
        gathered = block.view(1, rows, 128)
    out, offs = [None] * n_images, [0] * world
    for i in range(n_images):
        r = owner_of(i, world)
        out[i] = gathered[r, offs[r]:offs[r] + counts_h[i]]
        offs[r] += counts_h[i]
    return out, counts_h


def match_collection(local, n_images, ratio_threshold=0.75, ctx=None, matcher=None, group=None,
                     both_directions=False, gather_to_rank0=True, device=None):
    """All-pairs Lowe-ratio matching of a collection.  Returns {(i, j): (idx_i, idx_j, dist)};
    complete on rank 0 (when gather_to_rank0), this rank's share elsewhere."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if matcher is None:
        if ctx is None:
            raise ValueError("match_collection needs a SiftContext (there is no CPU matcher in the product)")
        matcher = lambda a, b: ctx.match(a, b, ratio_threshold)  # noqa: E731
    descs, counts = all_gather_descriptors(local, n_images, group=group, device=device)
    mine = partition_pairs(counts, world, both_directions)[rank]
    result = {}
    for (i, j) in mine:
        result[(i, j)] = matcher(descs[i], descs[j])
    if world > 1 and gather_to_rank0:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(result, parts, dst=0, group=group)
        if rank == 0:
            result = {}
            for p in parts:
                result.update(p)
    return result


# ------------------------------------------------------------------------------------------
# Collection driver: the stitching datasets shipped with the reference
# (stitching/collection/Dataset/*/) carry a "<name>-STITCH-GRAPH.txt" with lines such as
#   {images_count | 21 | images count}
#   {center_image_index | 4 | center image index}
#   {matching_graph_image_edges-3 | 4,14 | matching graph image edge 3}
# i.e. image 3 is to be matched against images 4 and 14.  The driver detects every image once and
# matches exactly the listed edges (or all pairs when no graph is given).
# ------------------------------------------------------------------------------------------
def parse_stitch_graph(text):
    """Returns (images_count, center_image_index, sorted list of (i, j) edges with i < j)."""
    import re
    count, center, edges = None, None, set()
    for m in re.finditer(r"\{\s*([A-Za-z_]+)(?:-(\d+))?\s*\|\s*([^|]*?)\s*\|[^}]*\}", text):
        key, idx, val = m.group(1), m.group(2), m.group(3)
        if key == "images_count":
            count = int(val)
        elif key == "center_image_index":
            center = int(val)
        elif key == "matching_graph_image_edges" and idx is not None:
            i = int(idx)
            for tok in val.split(","):
                tok = tok.strip()
                if tok:
                    j = int(tok)
                    if i != j:
                        edges.add((min(i, j), max(i, j)))
    if count is None:
        raise ValueError("no images_count entry in the stitch graph")
    for i, j in edges:
        if not (0 <= i < count and 0 <= j < count):
            raise ValueError(f"edge ({i}, {j}) outside 0..{count - 1}")
    return count, center, sorted(edges)


def match_graph(ctx, keypoints, edges=None, ratio_threshold=0.75):
    """keypoints: list of record arrays (one per image, from SiftContext.detect).  Returns
    {(i, j): (idx_i, idx_j, dist)} for the listed edges (all pairs if edges is None)."""
    n = len(keypoints)
    if edges is None:
        edges = pair_list(n)
    descs = [np.ascontiguousarray(k["desc"]) for k in keypoints]
    return {(i, j): ctx.match(descs[i], descs[j], ratio_threshold) for (i, j) in edges}


def run_dataset(directory, device=0, ratio_threshold=0.75, all_pairs=False, **detect_params):
    """Detect every image of a dataset directory and match its stitch graph.  Images are decoded
    with PIL -- fine for a driver, but NOT pixel-identical to the reference's stb decoder on JPEGs;
    parity runs feed stb-decoded pixels (tests/golden/make_golden.py)."""
    import glob
    import os
    from PIL import Image
    from .api import SiftContext

    graphs = sorted(glob.glob(os.path.join(directory, "*STITCH-GRAPH.txt")))
    files = sorted(f for f in glob.glob(os.path.join(directory, "*"))
                   if f.lower().endswith((".jpg", ".jpeg", ".png", ".bmp")))
    edges = None
    if graphs and not all_pairs:
        count, _, edges = parse_stitch_graph(open(graphs[0]).read())
        if count != len(files):
            raise ValueError(f"{graphs[0]} lists {count} images, the directory holds {len(files)}")
    images = [np.asarray(Image.open(f).convert("RGB")) for f in files]
    w = max(im.shape[1] for im in images)
    h = max(im.shape[0] for im in images)
    with SiftContext(w, h, device) as ctx:
        kps = [ctx.detect(im, **detect_params) for im in images]
        matches = match_graph(ctx, kps, edges, ratio_threshold)
    return files, kps, matches


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 2:
        sys.exit("usage: python -m sift_project_b200.collection <dataset directory> [--all-pairs]")
    names, kps_, res = run_dataset(sys.argv[1], all_pairs="--all-pairs" in sys.argv[2:])
    for f, k in zip(names, kps_):
        print(f"{f}: {len(k)} keypoints")
    for (i, j), (ia, ib, d) in sorted(res.items()):
        print(f"edge {i}-{j}: {len(ia)} matches")
