"""Generate the committed golden vectors from the REAL reference (oracle/_ref/libhohref.so, built
from /root/reference by oracle/Makefile).  Run in the development container only:

    python tests/golden/make_golden.py

Outputs tests/golden/*.npz.  The fixtures pin both the CPU oracle (-m "not gpu" tests) and the
CUDA path (-m gpu tests) to reference outputs on boxes where /root/reference does not exist.
Citations: entropy_encoding.hpp:8, prediction.hpp:6/46/153, unprediction.hpp:6,
layer_encode.hpp:11, choh.cpp:104/394, SURVEY.md §8(c)/(d).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

STOCK = [0x0001, 0x0002, 0x0020, 0x0010, 0xffbf, 0x0003, 0xfffd, 0xfffb, 0xfff7, 0xffef, 0xffdf,
         0xff7f, 0xfdff, 0xffff]


def planes_of(rgb):
    n = rgb.size // 3
    g, rg, bg = (np.zeros(n, np.uint16) for _ in range(3))
    ol.ref().ref_subtract_green(rgb, rgb.size, g, rg, bg)
    return g, rg, bg


def entropy_cases():
    rng = np.random.default_rng(1234)
    out = {}
    cases = []
    text = np.frombuffer(open("/root/reference/simple_entropy_encoder.cpp", "rb").read(), np.uint8)
    cases.append(("text817", text.astype(np.uint16), 256, 12))
    cases.append(("empty", np.zeros(0, np.uint16), 256, 10))
    cases.append(("single", np.full(300, 7, np.uint16), 256, 12))
    cases.append(("range1", np.zeros(49, np.uint16), 1, 8))
    cases.append(("predmap", rng.integers(0, 4, 49).astype(np.uint16), 4, 8))
    cases.append(("noise9", rng.integers(0, 512, 4000).astype(np.uint16), 512, 15))
    for pb in (12, 13, 14, 15, 16, 17, 18, 19):
        s = np.clip(np.rint(rng.laplace(256, 6.0, 20000)), 0, 511).astype(np.uint16)
        cases.append((f"lap9_pb{pb}", s, 512, pb))
    s = np.clip(np.rint(rng.laplace(128, 2.5, 65536)), 0, 255).astype(np.uint16)
    cases.append(("lap8_64k", s, 256, 15))
    s = np.clip(np.rint(rng.laplace(128, 30.0, 3000)), 0, 255).astype(np.uint16)
    cases.append(("wide8_mode1", s, 256, 8))
    geo = ol.synth_symbols(65536, 99).astype(np.uint16)
    cases.append(("geo64k", geo, 256, 12))
    for name, sym, rng_, pb in cases:
        stream = ol.ref_encode_entropy(sym, rng_, pb)
        out[f"{name}__sym"] = sym
        out[f"{name}__par"] = np.array([rng_, pb], np.int64)
        out[f"{name}__out"] = stream
    np.savez_compressed(os.path.join(HERE, "entropy.npz"), **out)
    print("entropy.npz", len(cases), "cases")


def predict_cases():
    rng = np.random.default_rng(77)
    R = ol.ref()
    out = {}
    k = 0
    for (w, h) in [(5, 4), (40, 40), (41, 83), (100, 64), (256, 256), (256, 270)]:
        rgb = ol.synth_rgb(w, h, 1 + k)
        g, rg, bg = planes_of(rgb)
        for depth, plane in ((8, g), (9, rg)):
            if k % 3 == 2:  # harsher content
                plane = rng.integers(0, 1 << depth, w * h).astype(np.uint16)
            xt, yt = (w + 39) // 40, (h + 39) // 40
            fast = np.zeros(w * h, np.uint16)
            R.ref_predict_fastpath(plane, plane.size, w, h, depth, fast)
            tm = rng.choice(STOCK, xt * yt).astype(np.uint16)
            allr = np.zeros(w * h, np.uint16)
            R.ref_predict_all(plane, plane.size, w, h, depth, xt, yt, tm, allr)
            sec = []
            for t in range(xt * yt):
                buf = np.zeros(1700, np.uint16)
                n = R.ref_predict_section(plane, plane.size, w, h, depth, xt, yt, t % xt, t // xt,
                                          int(tm[t]), buf)
                sec.append(buf[:n])
            un = np.zeros(w * h, np.uint16)
            R.ref_unpredict_all(np.concatenate([allr, np.zeros(4, np.uint16)]), allr.size, w, h,
                                depth, xt, yt, tm, np.zeros(w * h, np.uint16), un)
            assert (un == plane).all()
            tag = f"c{k}"
            out[f"{tag}__par"] = np.array([w, h, depth, xt, yt], np.int64)
            out[f"{tag}__plane"] = plane
            out[f"{tag}__fast"] = fast
            out[f"{tag}__map"] = tm
            out[f"{tag}__all"] = allr
            out[f"{tag}__sec"] = np.concatenate(sec)
            out[f"{tag}__secn"] = np.array([len(s) for s in sec], np.int64)
            k += 1
    np.savez_compressed(os.path.join(HERE, "predict.npz"), **out)
    print("predict.npz", k, "cases")


def layer_and_tile_cases():
    R = ol.ref()
    out = {}
    img54 = np.array([1, 0, 1, 1, 5, 1, 255, 1, 1, 1, 254, 1, 1, 1, 1, 1, 1, 14, 1, 1], np.uint16)
    out["l54__plane"] = img54
    out["l54__par"] = np.array([5, 4, 8, 2], np.int64)
    out["l54__out"] = ol.ref_layer_encode(img54, 5, 4, 8, 2)
    k = 0
    for (w, h, mode) in [(64, 48, 0), (100, 90, 1), (128, 128, 2), (100, 64, 3), (256, 256, 0),
                         (256, 256, 2), (256, 256, 4)]:
        rgb = ol.synth_rgb(w, h, 11 + k)
        g, rg, bg = planes_of(rgb)
        for depth, plane in ((8, g), (9, bg)):
            tag = f"l{k}"
            out[f"{tag}__plane"] = plane
            out[f"{tag}__par"] = np.array([w, h, depth, mode], np.int64)
            out[f"{tag}__out"] = ol.ref_layer_encode(plane, w, h, depth, mode)
            k += 1
    # encode_tile (choh.cpp:104): example.rgb and generator tiles
    ex = np.frombuffer(open("/root/reference/example.rgb", "rb").read(), np.uint8).copy()
    for mode in (0, 1):
        buf = np.zeros(4096, np.uint8)
        n = R.ref_encode_tile(ex, ex.size, buf, 2, 2, mode)
        out[f"tile_example_s{mode}"] = buf[:n].copy()
    for (w, h, mode, seed) in [(256, 256, 0, 1), (256, 256, 2, 1), (96, 80, 4, 5)]:
        rgb = ol.synth_rgb(w, h, seed)
        buf = np.zeros(rgb.size * 3 + 4096, np.uint8)
        n = R.ref_encode_tile(rgb, rgb.size, buf, w, h, mode)
        out[f"tile_{w}x{h}_s{mode}_seed{seed}"] = buf[:n].copy()
    # whole-tool files (choh.cpp:394): keep md5 + size only
    files = {}
    with tempfile.TemporaryDirectory() as td:
        for (w, h, mode) in [(512, 512, 0), (512, 512, 2), (512, 512, 4), (768, 512, 0)]:
            rgb = ol.synth_rgb(w, h, 1)
            ip, op = os.path.join(td, "i.rgb"), os.path.join(td, "o.hoh")
            rgb.tofile(ip)
            rc = R.ref_choh_main(ip.encode(), op.encode(), w, h, mode)
            data = open(op, "rb").read()
            files[f"{w}x{h}_s{mode}"] = (len(data), hashlib.md5(data).hexdigest(), rc)
            if (w, h, mode) == (512, 512, 0):
                out["file_512x512_s0"] = np.frombuffer(data, np.uint8).copy()
    out["files_keys"] = np.array(list(files.keys()))
    out["files_size"] = np.array([v[0] for v in files.values()], np.int64)
    out["files_md5"] = np.array([v[1] for v in files.values()])
    np.savez_compressed(os.path.join(HERE, "layer_tile.npz"), **out)
    print("layer_tile.npz", k, "layer cases;", files)


def lz_cases():
    """find_lz_rgb (lz.hpp:6) on images with matches: LZ bytes + nuke map from the real reference."""
    rng = np.random.default_rng(2024)
    out = {}
    k = 0
    for kind in ("flat", "pattern", "rows", "few", "photo"):
        for (w, h, distance, bonus) in [(64, 48, 6, 0), (61, 37, 10, 2), (96, 64, 12, 0), (80, 50, 14, 10)]:
            img = ol.lz_test_image(rng, w, h, kind)
            lz, nuke = ol.ref_find_lz_rgb(img, w, h, distance, bonus)
            out[f"z{k}__rgb"] = img.ravel()
            out[f"z{k}__par"] = np.array([w, h, distance, bonus], np.int64)
            out[f"z{k}__lz"] = lz
            out[f"z{k}__nuke"] = np.packbits(nuke)
            k += 1
    np.savez_compressed(os.path.join(HERE, "lz.npz"), **out)
    print("lz.npz", k, "cases")


def unequal_tile_files():
    """Whole `choh` files for image sizes whose tile grid has unequal tiles (choh.cpp:459-474): size + md5."""
    keys, sizes, md5s = [], [], []
    with tempfile.TemporaryDirectory() as td:
        for (w, h, mode) in [(601, 523, 0), (601, 523, 2), (1000, 700, 0), (530, 300, 1)]:
            rgb = ol.synth_rgb(w, h, 1)
            ip, op = os.path.join(td, "i.rgb"), os.path.join(td, "o.hoh")
            rgb.tofile(ip)
            ol.ref().ref_choh_main(ip.encode(), op.encode(), w, h, mode)
            data = open(op, "rb").read()
            keys.append(f"{w}x{h}_s{mode}")
            sizes.append(len(data))
            md5s.append(hashlib.md5(data).hexdigest())
    np.savez_compressed(os.path.join(HERE, "files_unequal_tiles.npz"), files_keys=np.array(keys),
                        files_size=np.array(sizes, np.int64), files_md5=np.array(md5s))


if __name__ == "__main__":
    assert ol.have_ref(), "needs /root/reference (development container)"
    entropy_cases()
    predict_cases()
    layer_and_tile_cases()
    lz_cases()
    unequal_tile_files()
