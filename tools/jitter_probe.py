#!/usr/bin/env python
"""Call-to-call spread of hoh_encode_images / hoh_decode_images on one batch (development tool): prints the device time
of every call.  usage: jitter_probe.py [images] [w] [h] [mode] [calls]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n, w, h, mode, calls = [int(x) for x in (sys.argv[1:] + ["8192", "256", "256", "4", "10"][len(sys.argv) - 1:])]
mod = bench._load("hohgpu", os.path.join(ROOT, "hoh-ans_b200", "host", "hohgpu.py"))
g = mod.HohGpu(0)
geo = g.tile_geometry(w, h)
n_tiles = n * geo.tiles_per_image
raw = n * w * h * 3
rgb = np.zeros(raw, np.uint8)
bench.fill_images(rgb, 1, n, w, h, os.cpu_count() or 1)
d_rgb = g.alloc(raw).upload(rgb)
cap = raw + raw // 2 + 8192 * n_tiles
d_packed, d_off, d_tiles = g.alloc(cap), g.alloc((n_tiles + 1) * 8), g.alloc(n_tiles * mod.TILE_DT.itemsize)
d_back, d_st = g.alloc(raw), g.alloc(n_tiles * 4)
enc_ms, dec_ms = [], []
for i in range(calls):
    g.timer_start(0)
    g._ck(g.lib.hoh_encode_images(g.ctx, d_rgb.ptr, n, w, h, mode, 24, d_packed.ptr, cap, d_off.ptr, d_tiles.ptr), "enc")
    g.timer_stop(0)
    enc_ms.append(round(g.timer_ms(0), 1))
    g.timer_start(1)
    g._ck(g.lib.hoh_decode_images(g.ctx, d_packed.ptr, cap, d_off.ptr, n, w, h, d_back.ptr, d_st.ptr), "dec")
    g.timer_stop(1)
    dec_ms.append(round(g.timer_ms(1), 1))
print(json.dumps({"images": n, "size": f"{w}x{h}", "mode": mode, "encode_ms": enc_ms, "decode_ms": dec_ms}))
