#!/usr/bin/env python
"""LZ match finder (find_lz_rgb, lz.hpp:6) throughput: GPU batch vs the reference on the host cores.

    python tools/bench_lz.py [--tiles 4096] [--size 256] [--distances 6,11] [--cpu-tiles 16]

Tiles are synthetic photos (SURVEY section 8(d) generator: no matches, the common case for `choh` on
photographic input) — the reference still compares every pixel with every candidate distance.  Prints one
JSON line per distance.  Development tool: bench.py stays the judged benchmark."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (synthetic image generator + module loader)


def _cpu(args):
    seed, w, h, distance = args
    import oracle_lib as ol
    rgb = ol.synth_rgb(w, h, seed)
    t0 = time.perf_counter()
    if ol.have_ref():
        ol.ref_find_lz_rgb(rgb, w, h, distance, 0)
    else:
        ol.orc_find_lz_rgb(rgb, w, distance, 0)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=4096)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--distances", default="6,11")
    ap.add_argument("--cpu-tiles", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    mod = bench._load("hohgpu", os.path.join(ROOT, "hoh-ans_b200", "host", "hohgpu.py"))
    g = mod.HohGpu(0)
    w = h = a.size
    n = a.tiles
    npx = w * h
    rgb = np.zeros(n * npx * 3, np.uint8)
    bench.fill_images(rgb, 1, n, w, h, os.cpu_count() or 1)
    stride = int(g.lib.hoh_find_lz_stride(w, h))
    d_rgb = g.alloc(rgb.nbytes).upload(rgb)
    d_nuke = g.alloc(n * npx)
    d_lz = g.alloc(n * stride)
    d_size = g.alloc(n * 4)
    d_st = g.alloc(n * 4)
    for distance in [int(x) for x in a.distances.split(",")]:
        def run():
            g._ck(g.lib.hoh_find_lz_rgb_batch(g.ctx, d_rgb.ptr, n, w, h, distance, 0, None, d_nuke.ptr, d_lz.ptr, stride,
                                              d_size.ptr, d_st.ptr), "hoh_find_lz_rgb_batch")
        run()
        g.sync()
        g.timer_start(0)
        for _ in range(a.steps):
            run()
        g.timer_stop(0)
        ms = g.timer_ms(0) / a.steps
        g.profile_begin()
        run()
        prof = g.profile_end()
        assert (d_st.download(np.int32, n) == 0).all()
        with ProcessPoolExecutor(os.cpu_count()) as ex:
            t0 = time.perf_counter()
            per = list(ex.map(_cpu, [(1 + i, w, h, distance) for i in range(a.cpu_tiles)]))
            wall = time.perf_counter() - t0
        print(json.dumps({
            "what": "find_lz_rgb", "tiles": n, "tile": f"{w}x{h}", "distance": distance,
            "gpu_ms": ms, "gpu_mbs": n * npx * 3 / (ms / 1e3) / 1e6,
            "pairs_per_s": n * npx * ((1 << distance) + (256 if distance > 8 else 0)) / (ms / 1e3),
            "kernels_ms": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]},
            "cpu_ref_ms_per_tile_one_core": 1e3 * float(np.mean(per)),
            "cpu_mbs_all_cores": a.cpu_tiles * npx * 3 / wall / 1e6, "cpu_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
