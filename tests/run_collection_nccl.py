"""torchrun script (N GPUs, NCCL): detect on image shards, then all-pairs matching twice -- through the Python
host layer (torch.distributed all-gather, collection.match_collection) and through the C ABI
(sift_b200_comm_attach + sift_b200_collection_match: NCCL communicator and ncclAllGather inside libsift_b200.so) --
and check both against the oracle on rank 0.  Used by tests/test_gpu_multi.py and directly:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_collection_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sift_project_b200 as S  # noqa: E402
from sift_project_b200 import collection as Cn  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_images = 6
    imgs = [O.synth_image(240, 320, seed=50 + i) for i in range(n_images)]
    ctx = S.SiftContext(320, 240, device=local)
    mine = {}
    kps = {}
    for i in range(n_images):
        if Cn.owner_of(i, world) == rank:
            k = ctx.detect(imgs[i])
            kps[i] = k
            mine[i] = torch.from_numpy(np.ascontiguousarray(k["desc"])).cuda()
    res = Cn.match_collection(mine, n_images, ctx=ctx)
    # ---- the same through the C ABI: the communicator lives in the library ----
    uid = [S.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_attach(uid[0], world, rank)
    assert ctx.comm_info() == (world, rank)
    owned = [mine[i] for i in sorted(mine)]
    n_pairs = ctx.collection_match(n_images, owned)
    pi, pj, rows = ctx.collection_pairs()
    assert n_pairs == len(pi)
    native = {(int(i), int(j)): ctx.collection_fetch(q, r) for q, (i, j, r) in enumerate(zip(pi, pj, rows))}
    digest = ctx.collection_digest(all_ranks=True)
    all_native = [None] * world
    dist.all_gather_object(all_native, native)
    all_digest = [None] * world
    dist.all_gather_object(all_digest, digest)
    # identical keypoints no matter which GPU detected them (bitwise: sharding is by image)
    all_k = [None] * world
    dist.all_gather_object(all_k, kps)
    if rank == 0:
        merged = {}
        for d in all_k:
            merged.update(d)
        solo = {i: ctx.detect(imgs[i]) for i in range(n_images)}
        for i in range(n_images):
            assert merged[i].tobytes() == solo[i].tobytes(), i
        assert sorted(res) == Cn.pair_list(n_images)
        total = 0
        for (i, j), (ia, ib, d) in res.items():
            wa, wb, wd = O.match(O.port(), solo[i]["desc"], solo[j]["desc"])
            assert np.array_equal(ia, wa) and np.array_equal(ib, wb) and np.array_equal(d, wd), (i, j)
            total += len(ia)
        merged_native = {}
        for d in all_native:
            assert not (set(d) & set(merged_native))      # a pair is never split or duplicated
            merged_native.update(d)
        assert sorted(merged_native) == Cn.pair_list(n_images)
        for key, (ia, ib, d) in merged_native.items():
            assert np.array_equal(ia, res[key][0]) and np.array_equal(ib, res[key][1]) and np.array_equal(d, res[key][2]), key
        assert len(set(all_digest)) == 1 and all_digest[0][0] == total      # all-reduced digest: same on every rank
        # ... and equal to the digest of the whole collection matched by ONE context
        solo_ctx = S.SiftContext(64, 64, device=local)
        solo_ctx.collection_match(n_images, [np.ascontiguousarray(solo[i]["desc"]) for i in range(n_images)])
        assert solo_ctx.collection_digest(all_ranks=False) == all_digest[0]
        solo_ctx.close()
        print(f"collection ok: world {world}, {n_images} images, {len(res)} pairs, {total} matches, "
              f"digest {all_digest[0][1]:016x} (C ABI path == torch path == oracle == 1-GPU digest)")
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
