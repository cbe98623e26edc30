"""encode_tile (choh.cpp:104-382) at cruncher modes 1-4 with every hot-path piece on the device and batched over
tiles: LZ match finder with the mode's seek window and the colour-count bonus (hoh_find_lz_rgb_batch), colour
planes (hoh_subtract_green / hoh_channel_picker), and the whole of layer_encode for all planes of all tiles in
one call with the NUKE maps (hoh_layer_encode_batch).  Host code only applies encode_tile's emit rules
(colour-mode comparison of sizes, channel-order byte, varints).  Result: the reference's tile bytes."""
import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu
DIST = {0: 6, 1: 10, 2: 11, 3: 12, 4: 14}


def gpu_encode_tiles(g, tiles, w, h, mode):
    n = len(tiles)
    flat = np.concatenate([t.ravel() for t in tiles])
    lz = g.find_lz_rgb_batch(flat, n, w, h, DIST[mode])
    nuke = np.concatenate([rec[1] for rec in lz])
    planes8, planes9 = [], []
    for t in tiles:
        gp, rg, bg = g.subtract_green(t.ravel())
        planes8.append(gp)
        planes9 += [rg, bg]
        if mode > 2:  # choh.cpp:263-290: plain R and B as alternatives
            planes8 += [g.channel_picker(t.ravel(), 3, 0), g.channel_picker(t.ravel(), 3, 2)]
    per8 = 3 if mode > 2 else 1
    enc8 = g.layer_encode_batch(np.concatenate(planes8), n * per8, w, h, 8, mode, nuke, per8)
    enc9 = g.layer_encode_batch(np.concatenate(planes9), n * 2, w, h, 9, mode, nuke, 2)
    out = []
    for i in range(n):
        assert lz[i][2] == 0
        sub_green = [enc8[i * per8][0], enc9[2 * i][0], enc9[2 * i + 1][0]]
        alt = [enc8[i * per8 + 1][0], enc8[i * per8 + 2][0]] if mode > 2 else None
        colour_mode, chans = ol.pick_colour_mode(sub_green, alt)
        out.append(ol.assemble_tile(lz[i][0].tobytes(), colour_mode, chans))
    return out


@pytest.mark.parametrize("w,h,mode", [(96, 80, 1), (100, 64, 2), (90, 70, 3), (64, 96, 4), (81, 41, 2)])
def test_encode_tile_modes_vs_oracle(w, h, mode):
    g = gpu_lib.gpu()
    rng = np.random.default_rng(100 * mode + w)
    tiles = [ol.photo_with_repeats(rng, w, h, 300 + 10 * mode + i) for i in range(4)]
    if mode > 2:  # one tile whose channels are uncorrelated, so that plain RGB (colour mode 2) wins
        tiles[1][..., 0] = rng.integers(0, 256, (h, w))
        tiles[1][..., 2] = rng.integers(0, 256, (h, w))
    got = gpu_encode_tiles(g, tiles, w, h, mode)
    colour_modes, nuked = set(), 0
    for i, t in enumerate(tiles):
        want, nuke = ol.orc_encode_tile_subgreen(t, mode)
        assert got[i] == want, (i, w, h, mode, len(got[i]), len(want))
        colour_modes.add(want[2])
        nuked += int(nuke.sum())
    assert nuked > 100
    assert (2 in colour_modes) if mode > 2 else colour_modes == {128}
