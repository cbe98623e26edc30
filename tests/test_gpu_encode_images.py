"""hoh_encode_images: encode_tile (choh.cpp:104-382) for every tile of a batch of images on the device, any
cruncher mode.  Checked against (a) the reference's own encode_tile outputs and whole `choh` files recorded in
tests/golden (bytes and md5), (b) the pinned oracle's tile encoder on images with LZ matches."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def _container():
    spec = importlib.util.spec_from_file_location(
        "hoh_container", os.path.join(HERE, "..", "hoh-ans_b200", "host", "container.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_encode_tile_golden():
    """ref_encode_tile outputs recorded from the real reference: 256x256 at -s0 and -s2, 96x80 at -s4."""
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "layer_tile.npz"))
    for (w, h, mode, seed) in [(256, 256, 0, 1), (256, 256, 2, 1), (96, 80, 4, 5)]:
        tiles, rec = g.encode_images(ol.synth_rgb(w, h, seed), 1, w, h, mode)
        assert rec["status"][0] == 0 and rec["flags"][0] == 0
        assert tiles[0] == z[f"tile_{w}x{h}_s{mode}_seed{seed}"].tobytes(), (w, h, mode)


@pytest.mark.parametrize("key", ["512x512_s0", "512x512_s2", "512x512_s4", "768x512_s0"])
def test_whole_choh_file_md5(key):
    """Stock `choh` files (choh.cpp:394-527, recorded size + md5): tiles from the device, container bytes from
    the host mirror of choh.cpp:436-506."""
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "layer_tile.npz"))
    i = list(z["files_keys"]).index(key)
    dims, mode = key.split("_s")
    w, h = (int(v) for v in dims.split("x"))
    tiles, rec = g.encode_images(ol.synth_rgb(w, h, 1), 1, w, h, int(mode))
    assert (rec["status"] == 0).all()
    geo = g.tile_geometry(w, h)
    data, printed = _container().assemble_file(w, h, geo.x_tiles, geo.y_tiles, tiles, rec["flags"])
    assert len(data) == int(z["files_size"][i])
    assert hashlib.md5(data).hexdigest() == str(z["files_md5"][i])
    if key == "512x512_s0":
        assert data == z["file_512x512_s0"].tobytes()


@pytest.mark.parametrize("w,h,mode,n", [(512, 512, 0, 3), (512, 256, 1, 2), (96, 80, 2, 5), (90, 70, 3, 4), (64, 96, 4, 3),
                                        (768, 540, 2, 1)])
def test_encode_images_vs_oracle_tiles(w, h, mode, n):
    """Batches of images with LZ matches (and, at modes 3-4, one whose plain-RGB alternative wins); 768x540 is cut
    into 256x270 tiles like the 4K frames of BASELINE config 3."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(31 * mode + w)
    imgs = [ol.photo_with_repeats(rng, w, h, 500 + 7 * mode + i) for i in range(n)]
    if mode > 2:
        imgs[1][..., 0] = rng.integers(0, 256, (h, w))
        imgs[1][..., 2] = rng.integers(0, 256, (h, w))
    tiles, rec = g.encode_images(np.concatenate([i.ravel() for i in imgs]), n, w, h, mode)
    assert (rec["status"] == 0).all()
    geo = g.tile_geometry(w, h)
    colour_modes, nuked = set(), 0
    for i, img in enumerate(imgs):
        for t in range(geo.tiles_per_image):
            x0, y0 = (t % geo.x_tiles) * geo.tile_w, (t // geo.x_tiles) * geo.tile_h
            tile = np.ascontiguousarray(img[y0:y0 + geo.tile_h, x0:x0 + geo.tile_w])
            want, nuke = ol.orc_encode_tile_subgreen(tile, mode)
            k = i * geo.tiles_per_image + t
            assert tiles[k] == want, (i, t, mode, len(tiles[k]), len(want))
            assert rec["size"][k] == len(want) and rec["colour_mode"][k] == want[2]
            colour_modes.add(want[2])
            nuked += int(nuke.sum())
    assert nuked > 100
    assert (2 in colour_modes) if mode > 2 else colour_modes == {128}


def test_flags_for_grey_and_palette_tiles():
    g = gpu_lib.gpu()
    w = h = 64
    rng = np.random.default_rng(3)
    photo = ol.synth_rgb(w, h, 9).reshape(h, w, 3)
    grey = np.repeat(photo[..., 1:2], 3, axis=2)
    pal = rng.integers(0, 256, (10, 3)).astype(np.uint8)[rng.integers(0, 10, (h, w))]
    tiles, rec = g.encode_images(np.concatenate([photo.ravel(), grey.ravel(), pal.ravel()]), 3, w, h, 0)
    assert list(rec["flags"]) == [0, 3, 2]  # a grey 8-bit tile also has at most 256 colours


@pytest.mark.parametrize("key", ["601x523_s0", "601x523_s2", "1000x700_s0", "530x300_s1"])
def test_whole_choh_file_md5_unequal_tiles(key):
    """Image sizes whose tile grid has unequal tiles (the last column / row takes what is left, choh.cpp:459-474:
    601x523 has four tile shapes): stock `choh` files by size and md5."""
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "files_unequal_tiles.npz"))
    i = list(z["files_keys"]).index(key)
    dims, mode = key.split("_s")
    w, h = (int(v) for v in dims.split("x"))
    tiles, rec = g.encode_images(ol.synth_rgb(w, h, 1), 1, w, h, int(mode))
    assert (rec["status"] == 0).all()
    geo = g.tile_geometry(w, h)
    data, printed = _container().assemble_file(w, h, geo.x_tiles, geo.y_tiles, tiles, rec["flags"])
    assert len(data) == int(z["files_size"][i])
    assert hashlib.md5(data).hexdigest() == str(z["files_md5"][i])


@pytest.mark.parametrize("mode", [0, 2])
def test_host_forms_round_trip_and_match_device_forms(mode):
    """hoh_encode_images_host / hoh_decode_images_host (what tools/choh_batch.cpp and dhoh_batch.cpp call): same
    tile bytes as the device-pointer form, exact round trip, two image shapes one after the other in one context
    (the cached scratch of the first shape is released, the staging of the running call is not)."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(5 + mode)
    for (w, h, n) in [(601, 523, 3), (256, 256, 5), (601, 523, 2)]:
        imgs = np.concatenate([ol.photo_with_repeats(rng, w, h, 70 + i).ravel() for i in range(n)])
        packed, off, rec = g.encode_images_host(imgs, n, w, h, mode, 24)
        assert (rec["status"] == 0).all()
        tiles, rec2 = g.encode_images(imgs, n, w, h, mode, 24)
        assert [packed[int(off[t]):int(off[t + 1])].tobytes() for t in range(len(tiles))] == tiles
        assert np.array_equal(rec["start"], off[:-1]) and np.array_equal(rec["size"], rec2["size"])
        back, st = g.decode_images_host(packed, off, n, w, h)
        assert (st == 0).all() and np.array_equal(back, imgs)


@pytest.mark.parametrize("w,h,mode,flags", [(256, 192, 2, 0), (256, 192, 2, 24), (128, 128, 4, 0), (128, 128, 4, 24),
                                            (96, 40, 1, 0)])
def test_candidates_picked_from_size_intervals(monkeypatch, w, h, mode, flags):
    """layer_encode's candidates (layer_encode.hpp:334-392) are compared through size intervals derived from their
    histograms and tables, and only the ones the outcome needs are coded (k_layer_plan).  The result must be the bytes
    of coding every candidate on the path, which HOH_LAYER_CODE_ALL=1 still does, with far fewer candidates coded."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(17)
    n = 6
    flat = np.concatenate([ol.photo_with_repeats(rng, w, h, 70 + i).ravel() for i in range(n - 1)] +
                          [ol.synth_rgb(w, h, 9)])
    g.layer_stats()
    tiles_a, rec_a = g.encode_images(flat, n, w, h, mode, flags)
    coded, unsettled, planes = g.layer_stats()
    monkeypatch.setenv("HOH_LAYER_CODE_ALL", "1")
    tiles_b, rec_b = g.encode_images(flat, n, w, h, mode, flags)
    coded_all, unsettled_all, planes_all = g.layer_stats()
    assert (rec_a["status"] == 0).all() and (rec_b["status"] == 0).all()
    assert tiles_a == tiles_b
    assert planes == planes_all > 0
    assert coded_all == 9 * planes_all and unsettled_all == planes_all
    assert coded <= 6 * unsettled + 2 * (planes - unsettled)
    assert unsettled * 2 <= planes  # the intervals settle most planes
