"""The N>1 path on CPU: two gloo ranks shard an image batch by index, code their shard with the
oracle standing in for the GPU, and the final gather + reductions reproduce the single-rank result."""
import multiprocessing as mp
import os
import sys

import numpy as np

import gpu_lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _shard_mod():
    return gpu_lib._load("hoh_shard", os.path.join(gpu_lib.PKG, "host", "shard.py"))


def _encode_sizes(first_image, count, w, h):
    """Per-channel payload sizes of images [first, first+count) (seed = 1 + image index)."""
    import oracle_lib as ol
    sizes = []
    for i in range(first_image, first_image + count):
        rgb = ol.synth_rgb(w, h, 1 + i)
        planes = [np.zeros(w * h, np.uint16) for _ in range(3)]
        ol.oracle().orc_subtract_green(rgb, rgb.size, *planes)
        for p, d in zip(planes, (8, 9, 9)):
            sizes.append(len(ol.orc_layer_encode(p, w, h, d, 0)[0]))
    return sizes


def _rank_main(rank, world_size, port, n_images, q):
    sys.path.insert(0, HERE)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world_size),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    sh = _shard_mod()
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    assert sh.world() == (rank, rank, world_size)
    lo, hi = sh.shard_range(n_images, rank, world_size)
    sizes = _encode_sizes(lo, hi - lo, 48, 40)
    dist.barrier()
    t_max = sh.reduce_max(10.0 + rank, world_size)       # device time: max over ranks
    total = sh.reduce_sum(float(sum(sizes)), world_size)  # bytes: sum over ranks
    tables = sh.gather_sizes(sizes, world_size)           # the final gather
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, t_max, total, tables))


def test_shard_ranges_cover_the_batch():
    sh = _shard_mod()
    for n in (1, 7, 256, 4096):
        for g in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, k, g) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    assert sh.weak_first_seed(4096, 0) == 1 and sh.weak_first_seed(4096, 3) == 1 + 3 * 4096


def test_two_rank_gloo_shard_and_gather():
    n_images, port = 5, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single = _encode_sizes(0, n_images, 48, 40)
    for rank, lo, hi, t_max, total, tables in res:
        assert t_max == 11.0                      # max over ranks of (10 + rank)
        assert total == float(sum(single))
        assert [s for t in tables for s in t] == single  # gather in rank order == unsharded order
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 2, 2, 5)
