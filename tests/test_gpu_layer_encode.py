"""Drop-in proof for cruncher modes >= 1: the HOST logic of layer_encode.hpp (header bytes, buffer
juggling, prob_bits search, the D7 stale-buffer behaviour — all of which stays host code) replayed here
line by line, with every hot-path call going to the GPU through the C-ABI shims
(hoh_channelpredict_fastpath, hoh_predictor_search, hoh_encode_entropy).  The assembled bytes must
equal what the real reference's layer_encode wrote (golden vectors) and what the oracle writes.
Also the reference's own layer_roundtrip_test (5x4 image, mode 2)."""
import os

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STOCK = [0x0001, 0x0002, 0x0020, 0x0010, 0xffbf, 0x0003, 0xfffd, 0xfffb, 0xfff7, 0xffef, 0xffdf,
         0xff7f, 0xfdff, 0xffff]


def gpu_layer_encode(g, plane, w, h, depth, mode, nuke=None):
    """layer_encode.hpp:11-412 with the hot-path calls on the GPU."""
    size = w * h
    nuke = np.zeros(size, np.uint8) if nuke is None else nuke
    out = bytearray()
    best_size = (depth * size + (depth * size) % 8 + 1024) // 8          # :22
    out.append(0x10)                                                     # :57
    resid = g.channelpredict_fastpath(plane, w, h, depth)                # :63-75 (section 1x1 mask 0x0010)
    dense = resid[nuke == 0]                                             # :93-99
    rng = 1 << depth
    kept = b""
    work, st = g.encode_entropy(dense, rng, 15)                          # :106
    assert st == 0
    if len(work) < best_size:                                            # :115-120
        best_size, kept = len(work), work.tobytes()
    xt, yt = (w + 39) // 40, (h + 39) // 40
    if mode and (xt > 1 or yt > 1):                                      # :126-319
        tile_map, index_list, resid = g.predictor_search(plane, w, h, depth, mode)
        out += bytes([xt - 1, yt - 1])                                   # :276-277
        used = sorted(set(int(i) for i in index_list))
        out.append(len(used))                                            # :291
        remap = {}
        for m in used:                                                   # :292-304
            out += bytes([STOCK[m] >> 8, STOCK[m] & 0xff])
            remap[m] = len(remap)
        idx = np.array([remap[int(i)] for i in index_list], np.uint16)
        stream, st = g.encode_entropy(idx, len(used), 8)                 # :308-317
        assert st == 0
        out += stream.tobytes()
    else:                                                                # :320-325
        out += bytes([0, 0, 0x00, 0x10])
    if mode:                                                             # :326-392
        dense = resid[nuke == 0]
        s16, _ = g.encode_entropy(dense, rng, 16)
        s15, _ = g.encode_entropy(dense, rng, 15)
        up = len(s16) < len(s15)
        first = len(s16) if up else len(s15)
        if first < best_size:
            best_size = first                                            # size updated, bytes NOT kept (D7)
        for k in range(3):
            bits = 17 + k if up else 14 - k
            s, _ = g.encode_entropy(dense, rng, bits)
            if len(s) < best_size:
                best_size, kept = len(s), s.tobytes()
    kept = kept + bytes(max(0, best_size - len(kept)))                   # :396-398 copies best_size bytes
    out += kept[:best_size]
    return bytes(out)


def _cases(npz, suffix):
    return sorted({k[: -len(suffix)] for k in npz.files if k.endswith(suffix)})


def test_layer_roundtrip_test_image():
    """layer_roundtrip_test.cpp:7-12: 5x4 image, depth 8, mode 2 -> the 29 known bytes."""
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "layer_tile.npz"))
    got = gpu_layer_encode(g, z["l54__plane"], 5, 4, 8, 2)
    assert got.hex() == "1000000010817f1400017f818084807f82807c7d82808080838" "08d7380"


def test_layer_encode_golden_all_modes():
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "layer_tile.npz"))
    seen_modes = set()
    for name in _cases(z, "__plane"):
        w, h, depth, mode = (int(v) for v in z[name + "__par"])
        got = gpu_layer_encode(g, z[name + "__plane"], w, h, depth, mode)
        want = z[name + "__out"].tobytes()
        if len(got) == len(want):  # D7 copies stale/uninitialised tail bytes only when lengths differ
            assert got == want, (name, w, h, depth, mode)
        else:
            raise AssertionError((name, len(got), len(want)))
        seen_modes.add(mode)
    assert seen_modes >= {0, 1, 2, 3, 4} or len(seen_modes) >= 3


def test_layer_encode_random_vs_oracle():
    g = gpu_lib.gpu()
    rng = np.random.default_rng(21)
    for it, (w, h, depth, mode) in enumerate([(96, 64, 8, 1), (100, 90, 9, 2), (128, 128, 8, 3), (64, 100, 9, 4)]):
        yy, xx = np.mgrid[0:h, 0:w]
        c = 1 << depth
        plane = np.clip(np.rint((np.sin(xx / 11.0) + np.cos(yy / 5.0)) * c / 5 + c / 2 + rng.normal(0, 2.5, (h, w))),
                        0, c - 1).astype(np.uint16).ravel()
        want, _ = ol.orc_layer_encode(plane, w, h, depth, mode)
        got = gpu_layer_encode(g, plane, w, h, depth, mode)
        assert got == want.tobytes(), (it, w, h, depth, mode)


def _smooth_planes(rng, n, w, h, depth):
    yy, xx = np.mgrid[0:h, 0:w]
    c = 1 << depth
    out = []
    for i in range(n):
        f = 7.0 + 3.0 * (i % 5)
        noise = rng.normal(0, 0.5 + (i % 4) * 2.0, (h, w))
        out.append(np.clip(np.rint((np.sin(xx / f) + np.cos(yy / (f / 2))) * c / 5 + c / 2 + noise), 0, c - 1)
                   .astype(np.uint16).ravel())
    return np.concatenate(out)


@pytest.mark.parametrize("w,h,depth,mode,n", [(96, 64, 8, 1, 5), (100, 90, 9, 2, 4), (128, 128, 8, 3, 3),
                                              (64, 100, 9, 4, 3), (33, 27, 8, 2, 6), (81, 41, 8, 0, 4),
                                              (41, 43, 9, 1, 3), (2, 90, 8, 2, 3), (1, 85, 9, 1, 2), (90, 2, 8, 2, 2),
                                              (45, 1, 8, 3, 2)])
def test_layer_encode_batch_vs_oracle(w, h, depth, mode, n):
    """hoh_layer_encode_batch: the whole of layer_encode.hpp:11-412 on the device for n planes at once,
    byte-identical to the oracle's layer_encode plane by plane (odd plane sizes exercise the padded
    symbol layout; 33x27 is a single predictor cell, 81x41 a 3x2 grid)."""
    g = gpu_lib.gpu()
    rng = np.random.default_rng(1000 * mode + w)
    planes = _smooth_planes(rng, n, w, h, depth)
    got = g.layer_encode_batch(planes, n, w, h, depth, mode)
    for i in range(n):
        want, _ = ol.orc_layer_encode(planes[i * w * h:(i + 1) * w * h], w, h, depth, mode)
        payload, st, kept = got[i]
        assert st == 0
        assert payload == want.tobytes(), (i, w, h, depth, mode, kept, len(payload), len(want))


def test_layer_encode_batch_golden():
    """The reference's own layer_encode outputs (golden vectors, modes 0-4) through the batched call."""
    g = gpu_lib.gpu()
    z = np.load(os.path.join(G, "layer_tile.npz"))
    for name in _cases(z, "__plane"):
        w, h, depth, mode = (int(v) for v in z[name + "__par"])
        plane = z[name + "__plane"]
        got = g.layer_encode_batch(np.concatenate([plane, plane]), 2, w, h, depth, mode)
        want = z[name + "__out"].tobytes()
        for payload, st, kept in got:
            assert st == 0 and payload == want, (name, w, h, depth, mode, kept)
