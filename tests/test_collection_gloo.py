"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: pair partitioning, the descriptor
all-gather, and the gather of match lists.  The matcher is the oracle here (a test double for the
tcgen05 kernel); the GPU suite checks the kernel itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from sift_project_b200 import collection as Cn


def test_partition_covers_every_pair_once_and_balances():
    counts = [20000] * 16 + [5000] * 7 + [1, 0, 300]
    for world in (1, 2, 3, 8):
        parts = Cn.partition_pairs(counts, world)
        flat = [p for r in parts for p in r]
        assert sorted(flat) == Cn.pair_list(len(counts))
        load = [sum(counts[i] * counts[j] for i, j in r) for r in parts]
        assert max(load) <= 1.05 * (sum(load) / world) + 20000 * 20000
    both = Cn.partition_pairs([3, 4, 5], 2, both_directions=True)
    assert sorted(p for r in both for p in r) == sorted(Cn.pair_list(3, True))
    assert Cn.partition_pairs(counts, 4) == Cn.partition_pairs(counts, 4)  # deterministic


def test_native_partition_equals_the_python_one():
    """sift_b200_partition_pairs (C++, what sift_b200_collection_match uses) deals exactly like
    collection.partition_pairs (pure host code: needs the built library, no GPU)."""
    import sift_project_b200 as S
    rng = np.random.default_rng(3)
    for n, world, both in ((7, 2, False), (16, 8, False), (9, 3, True), (1, 2, False), (0, 1, False), (40, 4, False)):
        counts = rng.integers(0, 5000, n).tolist()
        if n > 3:
            counts[2] = counts[1]          # equal costs: the tie-break must agree too
        py = Cn.partition_pairs(counts, world, both)
        for r in range(world):
            assert S.partition_pairs_native(counts, world, r, both) == py[r], (n, world, r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sets(n_images):
    rng = np.random.default_rng(0)
    return [O.synth_descriptors(int(rng.integers(1, 90)) if i != 2 else 0, seed=100 + i) for i in range(n_images)]


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sets = _sets(n_images)
        local = {i: torch.from_numpy(sets[i]) for i in range(n_images) if Cn.owner_of(i, world) == rank}
        descs, counts = Cn.all_gather_descriptors(local, n_images)
        assert counts == [len(s) for s in sets]
        for i in range(n_images):
            assert np.array_equal(descs[i].numpy(), sets[i]), i
        matcher = lambda a, b: O.match(O.port(), a.numpy(), b.numpy())  # noqa: E731
        res = Cn.match_collection(local, n_images, matcher=matcher)
        if rank == 0:
            q.put({k: tuple(x.tolist() for x in v) for k, v in res.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 1])   # 1 image: rank 1 owns nothing
def test_world2_collection_matches_single_process(n_images):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sets = _sets(n_images)
    assert sorted(got) == Cn.pair_list(n_images)
    for (i, j), (ia, ib, d) in got.items():
        wa, wb, wd = O.match(O.port(), sets[i], sets[j])
        assert ia == wa.tolist() and ib == wb.tolist() and d == wd.tolist(), (i, j)


def test_single_process_path_needs_a_context():
    with pytest.raises(ValueError):
        Cn.match_collection({0: torch.zeros((3, 128), dtype=torch.uint8)}, 1)


def test_parse_stitch_graph():
    text = """{center_image_index | 2 | center image index}
{center_image_rotation_angle | 0 | center image rotation angle}
{images_count | 5 | images count}
{matching_graph_image_edges-0 | 1,4 | matching graph image edge 0}
{matching_graph_image_edges-1 | 2 | matching graph image edge 1}
{matching_graph_image_edges-3 | 4,2 | matching graph image edge 3}
"""
    count, center, edges = Cn.parse_stitch_graph(text)
    assert (count, center) == (5, 2)
    assert edges == [(0, 1), (0, 4), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        Cn.parse_stitch_graph("{matching_graph_image_edges-0 | 1 | x}")
    with pytest.raises(ValueError):
        Cn.parse_stitch_graph("{images_count | 2 | n}{matching_graph_image_edges-0 | 5 | x}")
