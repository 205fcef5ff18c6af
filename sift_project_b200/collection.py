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
        device = next(iter(local.values())).device if local else torch.device("cpu")
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
                     both_directions=False, gather_to_rank0=True):
    """All-pairs Lowe-ratio matching of a collection.  Returns {(i, j): (idx_i, idx_j, dist)};
    complete on rank 0 (when gather_to_rank0), this rank's share elsewhere."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if matcher is None:
        if ctx is None:
            raise ValueError("match_collection needs a SiftContext (there is no CPU matcher in the product)")
        matcher = lambda a, b: ctx.match(a, b, ratio_threshold)  # noqa: E731
    descs, counts = all_gather_descriptors(local, n_images, group=group)
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
