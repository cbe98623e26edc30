"""Host mirror of the .hoh container writer (hoh-ans_b200/host/container.py, choh.cpp:436-506): CPU only.
The tiles are cut out of a file the real `choh -s0` wrote (tests/golden) and the writer must put the same
file back together; untiled images reproduce the header-only file of SURVEY D1."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def _container():
    spec = importlib.util.spec_from_file_location(
        "hoh_container", os.path.join(HERE, "..", "hoh-ans_b200", "host", "container.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_varints_match_the_oracle():
    import oracle_lib as ol
    c = _container()
    for v in [0, 1, 127, 128, 300, 16383, 16384, 99999, (1 << 21) - 1, 1 << 21, 5000000]:
        assert c.write_varint(v) == ol.varint_bytes(v), v


def test_stock_file_is_reassembled_from_its_tiles():
    c = _container()
    data = np.load(os.path.join(G, "layer_tile.npz"))["file_512x512_s0"].tobytes()
    assert data.startswith(c.file_header(512, 512) + bytes([1, 1]))
    pos = len(c.file_header(512, 512)) + 2
    sizes = []
    for _ in range(3):
        b0, b1, b2 = data[pos], data[pos + 1], data[pos + 2]
        assert b0 & 0x80 and b1 & 0x80                      # tiles of ~100 KB: three-byte varints
        sizes.append(((b0 & 0x7f) << 14) + ((b1 & 0x7f) << 7) + b2)
        pos += 3
    tiles = []
    for s in sizes:
        tiles.append(data[pos:pos + s])
        pos += s
    tiles.append(data[pos:])
    out, printed = c.assemble_file(512, 512, 2, 2, tiles)
    assert out == data and printed == len(data)


def test_untiled_image_writes_only_the_header():
    c = _container()
    out, printed = c.assemble_file(2, 2, 1, 1, [bytes(26)])
    assert out == bytes([0x99, 0x48, 0x4f, 0x48, 0x02, 0x08, 0x01, 0x01])   # example.rgb at -s0
    assert printed == 34


def test_flagged_tiles_are_refused():
    """A tile the device emitted in subtract-green mode where the reference would have gone greyscale / indexed must not
    end up in a file silently (hoh_tile_result.flags)."""
    import pytest
    c = _container()
    tiles = [bytes(30)] * 4
    c.assemble_file(512, 512, 2, 2, tiles, [0, 0, 0, 0])
    with pytest.raises(ValueError):
        c.assemble_file(512, 512, 2, 2, tiles, [0, 2, 0, 0])
