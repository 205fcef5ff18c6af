"""torchrun script (N GPUs, NCCL): detect on image shards, all-gather the descriptors, all-pairs
match with the tcgen05 kernel, and check rank 0's result against the oracle.  Used by
tests/test_gpu_multi.py and directly:
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
        print(f"collection ok: world {world}, {n_images} images, {len(res)} pairs, {total} matches")
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
