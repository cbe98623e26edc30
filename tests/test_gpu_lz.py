"""LZ match finder on the device (find_lz_rgb, lz.hpp:6 — SURVEY §8(f) row 1) against the reference's golden
vectors and the pinned oracle: LZ record bytes and NUKE maps, per call and batched."""
import os

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_find_lz_rgb_golden_vectors():
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "lz.npz"))
    names = sorted({k[:-5] for k in z.files if k.endswith("__rgb")})
    assert len(names) == 20
    for name in names:
        w, h, distance, bonus = (int(v) for v in z[name + "__par"])
        lz, nuke = g.find_lz_rgb(z[name + "__rgb"], w, h, distance, bonus)
        assert np.array_equal(nuke, np.unpackbits(z[name + "__nuke"])[: w * h]), name
        assert np.array_equal(lz, z[name + "__lz"]), name


@pytest.mark.parametrize("w,h,distance", [(64, 48, 6), (100, 37, 10), (256, 256, 6), (256, 270, 11), (33, 700, 14),
                                          (300, 7, 12)])
def test_find_lz_rgb_batch_vs_oracle(w, h, distance):
    """Every kind of content in one batch (flat, periodic, repeated rows, few colours, photo); bonus derived on
    the device from the colour count (choh.cpp:138-154) must equal the oracle's."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(w * 1000 + h)
    kinds = ["flat", "pattern", "rows", "few", "photo", "few", "pattern"]
    tiles = [ol.lz_test_image(rng, w, h, k) for k in kinds]
    got = g.find_lz_rgb_batch(np.concatenate([t.ravel() for t in tiles]), len(tiles), w, h, distance)
    matches = 0
    for i, t in enumerate(tiles):
        mode = {6: 0, 10: 1, 11: 2, 12: 3, 14: 4}[distance]
        d, bonus = ol.orc_lz_params(t, mode)
        assert d == distance
        want, want_nuke, side = ol.orc_find_lz_rgb(t, w, distance, bonus)
        lz, nuke, st = got[i]
        assert st == 0
        assert np.array_equal(nuke, want_nuke), (i, kinds[i])
        assert np.array_equal(lz, want), (i, kinds[i], len(lz), len(want))
        matches += len(side[1])
    assert matches > 50


def test_find_lz_rgb_synthetic_photo_has_constant_record():
    """the generator's images have no matches, so the LZ record is the constant the survey's dumps
    show (tag + one 255 per 255 pixels + two empty streams) and NUKE is all zero."""
    g = gpu_lib.gpu()
    w = h = 256
    tiles = np.concatenate([ol.synth_rgb(w, h, 1 + i) for i in range(6)])
    got = g.find_lz_rgb_batch(tiles, 6, w, h, 6)
    want, want_nuke, _ = ol.orc_find_lz_rgb(tiles[: w * h * 3], w, 6, 0)
    for lz, nuke, st in got:
        assert st == 0 and not nuke.any()
        assert np.array_equal(lz, want)


def _varint(v):
    if v < 128:
        return bytes([v])
    if v < (1 << 14):
        return bytes([0x80 | (v >> 7), v & 0x7f])
    return bytes([0x80 | (v >> 14), 0x80 | ((v >> 7) & 0x7f), v & 0x7f])


@pytest.mark.parametrize("w,h", [(256, 256), (512, 512), (200, 120), (600, 512)])
def test_encode_tile_s0_with_lz_matches_reference_bytes(w, h):
    """encode_tile (choh.cpp:104-382) at -s0 for photographic tiles with LZ matches: the tile bytes assembled by
    the host rules of choh.cpp:112-116, 328-363 from the device's LZ record, NUKE-compacted channel payloads and
    sizes equal the reference's (the oracle's tile encoder = find_lz_rgb + subtract_green + 3 x layer_encode)."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(w + h)
    n_images = 3
    imgs = [ol.photo_with_repeats(rng, w, h, 40 + i) for i in range(n_images)]
    lz, packed, off, res = g.encode_tiles_s0_lz(np.concatenate([i.ravel() for i in imgs]), n_images, w, h)
    assert (res["status"] == 0).all()
    geo = g.tile_geometry(w, h)
    nuked = 0
    for i, img in enumerate(imgs):
        for t in range(geo.tiles_per_image):
            x0, y0 = (t % geo.x_tiles) * geo.tile_w, (t // geo.x_tiles) * geo.tile_h
            tile = np.ascontiguousarray(img[y0:y0 + geo.tile_h, x0:x0 + geo.tile_w])
            want, nuke = ol.orc_encode_tile_subgreen(tile, 0)
            nuked += int(nuke.sum())
            s = (i * geo.tiles_per_image + t) * 3
            got_ch = [packed[int(off[s + c]):int(off[s + c + 1])].tobytes() for c in range(3)]
            got = (bytes([0, 0, 128]) + lz[i * geo.tiles_per_image + t].tobytes() + bytes([0b00100100]) +
                   _varint(len(got_ch[0])) + _varint(len(got_ch[1])) + b"".join(got_ch))
            assert got == want, (i, t, len(got), len(want))
    assert nuked > 100
