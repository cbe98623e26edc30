"""BASELINE.json's configurations at (or near) their stated sizes: exact encode -> decode round trips, sums of
sizes, and byte parity against the oracle.  Config 2 is run at full size and EVERY one of its 49 152 channel payloads
is compared with the oracle's layer_encode bytes (the oracle runs on all host cores: ctypes releases the GIL);
configs 3 and 5 run on slices whose tiles have exactly the configurations' shapes (256x270 tiles of 4K frames at
-s2; untiled 256x256 thumbnails at -s4) with 64+ tiles each compared with the oracle's encode_tile bytes; config 4
at 2^26 symbols is in test_gpu_parity.py (test_rans_static_sweep_64mb_round_trip)."""
from concurrent.futures import ThreadPoolExecutor
import ctypes as C
import importlib.util
import os
import threading

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX_ENCODER = 24


def _bench():
    spec = importlib.util.spec_from_file_location("hoh_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _images(n, w, h, first_seed=1):
    """The section 8(d) generator, image i seeded 1 + i (tools/synth_gen.c is the C twin of orc_synth_rgb)."""
    rgb = np.zeros(n * w * h * 3, np.uint8)
    _bench().fill_images(rgb, first_seed, n, w, h, os.cpu_count() or 4)
    assert np.array_equal(rgb[: w * h * 3], ol.synth_rgb(w, h, first_seed))
    return rgb


def test_config2_full_size_roundtrip_and_sampled_parity():
    """4096 images of 512x512 at -s0: 49 152 streams through the fused mode-0 codec, decode == input, and the
    channel payloads of sampled tiles equal the oracle's layer_encode bytes."""
    g = gpu_lib.gpu()
    mod = gpu_lib.hohgpu()
    n, w, h = 4096, 512, 512
    rgb = _images(n, w, h)
    geo = g.tile_geometry(w, h)
    n_streams = n * geo.streams_per_image
    out_bytes = int(g.lib.hoh_encode_images_out_bytes(C.byref(geo), n))
    packed_cap = rgb.nbytes + rgb.nbytes // 4
    bufs = [g.alloc(rgb.nbytes).upload(rgb), g.alloc(out_bytes), g.alloc(n_streams * mod.RESULT_DT.itemsize),
            g.alloc(packed_cap), g.alloc((n_streams + 1) * 8), g.alloc(rgb.nbytes), g.alloc(n_streams * 4)]
    d_rgb, d_out, d_res, d_packed, d_off, d_back, d_st = bufs
    try:
        g._ck(g.lib.hoh_encode_images_s0(g.ctx, d_rgb.ptr, n, w, h, None, d_out.ptr, out_bytes, d_res.ptr, d_packed.ptr,
                                         packed_cap, d_off.ptr), "encode")
        g._ck(g.lib.hoh_decode_images_s0(g.ctx, d_packed.ptr, packed_cap, d_off.ptr, n, w, h, None, d_back.ptr, d_st.ptr),
              "decode")
        off = d_off.download(np.uint64, n_streams + 1)
        res = d_res.download(mod.RESULT_DT, n_streams)
        assert (res["status"] == 0).all() and (d_st.download(np.int32, n_streams) == 0).all()
        assert int(off[-1]) == int(res["size"].astype(np.uint64).sum())           # a checksum of the size table
        assert 0.50 < int(off[-1]) / rgb.nbytes < 0.54                           # section 8(d): about 0.52 at -s0
        assert np.array_equal(d_back.download(np.uint8, rgb.nbytes), rgb)
        packed = d_packed.download(np.uint8, int(off[-1]))
        ol.oracle()                                                             # built / loaded before the threads start

        def check_image(image):                                                 # all 12 channel payloads of one image
            img = rgb[image * w * h * 3:(image + 1) * w * h * 3].reshape(h, w, 3)
            bad = []
            for t in range(geo.tiles_per_image):
                x0, y0 = (t % geo.x_tiles) * geo.tile_w, (t // geo.x_tiles) * geo.tile_h
                planes = ol.orc_subtract_green(np.ascontiguousarray(img[y0:y0 + geo.tile_h, x0:x0 + geo.tile_w]))
                s = (image * geo.tiles_per_image + t) * 3
                for c, (p, d) in enumerate(zip(planes, (8, 9, 9))):
                    want, _ = ol.orc_layer_encode(p, geo.tile_w, geo.tile_h, d, 0)
                    if packed[int(off[s + c]):int(off[s + c + 1])].tobytes() != want.tobytes():
                        bad.append(s + c)
            return bad

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
            mismatches = [s for bad in pool.map(check_image, range(n)) for s in bad]
        assert mismatches == [], mismatches[:10]                                # 49 152 of 49 152 streams byte-exact
    finally:
        for b in bufs:
            b.free()


def test_config3_slice_4k_frames_mode2():
    """8 frames of 3840x2160 at -s2 (960 tiles of 256x270): exact round trip of the decodable variant, and the
    reference-format bytes of two sampled tiles equal the oracle's encode_tile."""
    g = gpu_lib.gpu()
    n, w, h = 8, 3840, 2160
    rgb = _images(n, w, h, 11)
    tiles, rec = g.encode_images(rgb, n, w, h, 2, FIX_ENCODER)
    assert (rec["status"] == 0).all() and len(tiles) == 960
    back, st = g.decode_images(tiles, n, w, h)
    assert (st == 0).all() and np.array_equal(back, rgb)
    ref_tiles, rec0 = g.encode_images(rgb[: w * h * 3], 1, w, h, 2, 0)        # stock bytes of frame 0: 120 tiles
    img = rgb[: w * h * 3].reshape(h, w, 3)
    ol.oracle()

    def check_tile(t):
        x0, y0 = (t % 15) * 256, (t // 15) * 270
        want, _ = ol.orc_encode_tile_subgreen(np.ascontiguousarray(img[y0:y0 + 270, x0:x0 + 256]), 2)
        return ref_tiles[t] == want

    picked = list(range(0, 120, 2)) + [111, 113, 115, 117, 119]                # 65 tiles incl. both edges of the grid
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        ok = list(pool.map(check_tile, picked))
    assert all(ok), [t for t, v in zip(picked, ok) if not v]


def test_config5_slice_thumbnails_mode4():
    """2048 untiled 256x256 thumbnails at -s4 (full predictor search, 2^14-pixel LZ window, RGB alternative):
    exact round trip, and the reference-format bytes of one sampled thumbnail equal the oracle's encode_tile."""
    g = gpu_lib.gpu()
    n, w, h = 2048, 256, 256
    rgb = _images(n, w, h, 101)
    tiles, rec = g.encode_images(rgb, n, w, h, 4, FIX_ENCODER)
    assert (rec["status"] == 0).all()
    back, st = g.decode_images(tiles, n, w, h)
    assert (st == 0).all() and np.array_equal(back, rgb)
    picked = list(range(3, n, 32))                                             # 64 thumbnails, stock bytes (flags = 0)
    sub = np.concatenate([rgb[k * w * h * 3:(k + 1) * w * h * 3] for k in picked])
    ref_tiles, _ = g.encode_images(sub, len(picked), w, h, 4, 0)
    ol.oracle()

    def check_thumb(i):
        want, _ = ol.orc_encode_tile_subgreen(sub[i * w * h * 3:(i + 1) * w * h * 3].reshape(h, w, 3), 4)
        return ref_tiles[i] == want

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        ok = list(pool.map(check_thumb, range(len(picked))))
    assert all(ok), [picked[i] for i, v in enumerate(ok) if not v]
