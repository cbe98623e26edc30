#!/usr/bin/env python
"""Whole-tile encoder (`hoh_encode_images` = encode_tile choh.cpp:104 for every tile) at any cruncher mode:
GPU batch vs the reference's encode_tile on the host cores.

    python tools/bench_modes.py [--images 16] [--width 3840] [--height 2160] [--modes 0,2] [--cpu-tiles 16]

BASELINE config 3 is 256 frames of 3840x2160 at mode 2; the default here is a 16-frame slice of it (1 920
tiles of 256x270), which already fills the machine because the kernels are bound by the serial chain of a
stream, not by the number of streams.  Prints one JSON line per mode.  Development tool: bench.py stays the
judged benchmark."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def _cpu(args):
    seed, w, h, mode = args
    import oracle_lib as ol
    rgb = ol.synth_rgb(w, h, seed)
    t0 = time.perf_counter()
    if ol.have_ref():
        buf = np.zeros(rgb.size * 3 + 4096, np.uint8)
        ol.ref().ref_encode_tile(rgb, rgb.size, buf, w, h, mode)
    else:
        ol.orc_encode_tile_subgreen(rgb.reshape(h, w, 3), mode)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--modes", default="0,2")
    ap.add_argument("--cpu-tiles", type=int, default=16)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--flags", type=int, default=24, help="encoder flags: 24 = HOH_FIX_ENCODER (decodable), 0 = stock bytes")
    a = ap.parse_args()
    mod = bench._load("hohgpu", os.path.join(ROOT, "hoh-ans_b200", "host", "hohgpu.py"))
    g = mod.HohGpu(0)
    w, h, n = a.width, a.height, a.images
    geo = g.tile_geometry(w, h)
    n_tiles = n * geo.tiles_per_image
    raw = n * w * h * 3
    rgb = np.zeros(raw, np.uint8)
    bench.fill_images(rgb, 1, n, w, h, os.cpu_count() or 1)
    d_rgb = g.alloc(raw).upload(rgb)
    packed_cap = raw + raw // 2 + 8192 * n_tiles
    d_packed = g.alloc(packed_cap)
    d_off = g.alloc((n_tiles + 1) * 8)
    d_tiles = g.alloc(n_tiles * mod.TILE_DT.itemsize)
    for mode in [int(x) for x in a.modes.split(",")]:
        def run():
            g._ck(g.lib.hoh_encode_images(g.ctx, d_rgb.ptr, n, w, h, mode, a.flags, d_packed.ptr, packed_cap, d_off.ptr,
                                          d_tiles.ptr), "hoh_encode_images")
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
        rec = d_tiles.download(mod.TILE_DT, n_tiles)
        total = int(d_off.download(np.uint64, n_tiles + 1)[-1])
        assert (rec["status"] == 0).all()
        # decode what was just written (hoh_decode_images) and check the round trip when the output is decodable
        d_back = g.alloc(raw)
        d_st = g.alloc(n_tiles * 4)

        def run_dec():
            g._ck(g.lib.hoh_decode_images(g.ctx, d_packed.ptr, packed_cap, d_off.ptr, n, w, h, d_back.ptr, d_st.ptr),
                  "hoh_decode_images")
        run_dec()
        g.sync()
        g.timer_start(1)
        for _ in range(a.steps):
            run_dec()
        g.timer_stop(1)
        dec_ms = g.timer_ms(1) / a.steps
        g.profile_begin()
        run_dec()
        dprof = g.profile_end()
        dec_ok = bool((d_st.download(np.int32, n_tiles) == 0).all())
        same = bool(np.array_equal(d_back.download(np.uint8, raw), rgb)) if dec_ok else False
        d_back.free()
        d_st.free()
        with ProcessPoolExecutor(os.cpu_count()) as ex:
            list(ex.map(_cpu, [(1, 32, 32, 0)] * (os.cpu_count() or 1)))  # start the workers, load the libraries
            t0 = time.perf_counter()
            per = list(ex.map(_cpu, [(1 + i, geo.tile_w, geo.tile_h, mode) for i in range(a.cpu_tiles)]))
            wall = time.perf_counter() - t0
        tile_raw = geo.tile_w * geo.tile_h * 3
        print(json.dumps({
            "what": "encode_tile for every tile (LZ + colour planes + layer_encode + emission)", "mode": mode,
            "images": n, "size": f"{w}x{h}", "tiles": n_tiles, "tile": f"{geo.tile_w}x{geo.tile_h}",
            "gpu_ms": ms, "gpu_raw_mbs": raw / (ms / 1e3) / 1e6, "compressed_ratio": total / raw,
            "colour_modes": {int(k): int(v) for k, v in zip(*np.unique(rec["colour_mode"], return_counts=True))},
            "kernels_ms": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:10]},
            "launches": int(sum(v[1] for v in prof.values())),
            "decode_ms": dec_ms, "decode_raw_mbs": raw / (dec_ms / 1e3) / 1e6, "decode_status_ok": dec_ok,
            "roundtrip_exact": same, "encoder_flags": a.flags,
            "decode_kernels_ms": {k: round(v[0], 3) for k, v in sorted(dprof.items(), key=lambda kv: -kv[1][0])[:6]},
            "cpu_ref_ms_per_tile_one_core": 1e3 * float(np.mean(per)),
            "cpu_raw_mbs_all_cores": a.cpu_tiles * tile_raw / wall / 1e6, "cpu_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
