"""The drop-in boundary RUNS: the reference's own host programs relinked against libhohgpu.so through
include/hoh_gpu_shim.hpp (dropin/_bin/*_gpu, built from /root/reference by dropin/Makefile and shipped prebuilt)
produce stock `choh`'s files byte for byte and pass the reference's own round-trip tests with every hot-path call
on the GPU; the repo's C++ batch writer / reader (tools/choh_batch.cpp, tools/dhoh_batch.cpp) write the same files
and read them back."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "dropin", "_bin")
G = os.path.join(HERE, "golden")
EXAMPLE_RGB = bytes([0xff, 0, 0, 0, 0xff, 0, 0xff, 0xff, 0, 0, 0, 0xff])  # the reference's example.rgb (2x2)


def _exe(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        if name.endswith("_gpu"):
            pytest.skip(f"{name} not prebuilt (needs the reference sources at build time: make -C dropin ref)")
        subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "tools"], check=True, capture_output=True)
    return path


def _golden(key):
    for f in ("layer_tile.npz", "files_unequal_tiles.npz"):
        z = np.load(os.path.join(G, f))
        keys = list(z["files_keys"])
        if key in keys:
            i = keys.index(key)
            return int(z["files_size"][i]), str(z["files_md5"][i])
    raise KeyError(key)


def _run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


# ---------------------------------------------------------------------------------------------------
# the reference's programs, relinked
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["512x512_s0", "512x512_s2", "512x512_s4", "601x523_s2"])
def test_choh_gpu_writes_stock_choh_files(tmp_path, key):
    """`choh_gpu in.rgb out.hoh W H -sN`: stock choh.cpp + layer_encode.hpp + lz.hpp, hot path on the GPU, file md5
    equal to the stock CPU build's (78c14f89..., beba681c..., d5ded2c1... for 512x512)."""
    dims, mode = key.split("_s")
    w, h = (int(v) for v in dims.split("x"))
    size, md5 = _golden(key)
    (tmp_path / "in.rgb").write_bytes(ol.synth_rgb(w, h, 1).tobytes())
    r = _run([_exe("choh_gpu"), str(tmp_path / "in.rgb"), str(tmp_path / "out.hoh"), str(w), str(h), f"-s{mode}"])
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "out.hoh").read_bytes()
    assert r.stdout.strip().splitlines()[-1] == str(size)
    assert len(data) == size and hashlib.md5(data).hexdigest() == md5


def test_choh_gpu_example_rgb(tmp_path):
    """BASELINE config 1: example.rgb 2x2 at -s0 prints 34 and writes the 8-byte header."""
    (tmp_path / "example.rgb").write_bytes(EXAMPLE_RGB)
    r = _run([_exe("choh_gpu"), str(tmp_path / "example.rgb"), str(tmp_path / "o.hoh"), "2", "2", "-s0"])
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "34"
    assert (tmp_path / "o.hoh").read_bytes() == bytes([0x99, 0x48, 0x4f, 0x48, 0x02, 0x08, 0x01, 0x01])
    r = _run([_exe("choh_gpu"), str(tmp_path / "example.rgb"), str(tmp_path / "o.hoh"), "2", "2", "-s1"])
    assert r.stdout.strip() == "37"


def test_reference_entropy_roundtrip_test_on_the_gpu(tmp_path):
    """entropy_roundtrip_test.sh with the relinked programs: the 817-byte known answer (656 B, md5 c5d8fa19...)
    and the decoder giving the input back."""
    z = np.load(os.path.join(G, "entropy.npz"))
    src = z["text817__sym"].astype(np.uint8).tobytes()
    assert hashlib.md5(src).hexdigest() == "183756103eab4031a94038552ae46be1"
    (tmp_path / "in").write_bytes(src)
    r = _run([_exe("simple_entropy_encoder_gpu"), str(tmp_path / "in"), str(tmp_path / "comp")])
    assert r.returncode == 0, r.stderr
    comp = (tmp_path / "comp").read_bytes()
    assert len(comp) == 656 and hashlib.md5(comp).hexdigest() == "c5d8fa190d78049ed26a5ebcf06e3ddd"
    assert comp == z["text817__out"].tobytes()
    r = _run([_exe("simple_entropy_decoder_gpu"), str(tmp_path / "comp"), str(tmp_path / "back")])
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "back").read_bytes() == src


def test_reference_layer_roundtrip_test_on_the_gpu():
    """layer_roundtrip_test.cpp relinked: layer_encode (mode 2) -> decode_layer through the shim."""
    r = _run([_exe("layer_roundtrip_gpu")])
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "Layer roundtrip: OK" in r.stdout


def test_dhoh_gpu_runs():
    r = _run([_exe("dhoh_gpu"), "--version"])
    assert r.returncode == 0 and "experimental software" in r.stdout


# ---------------------------------------------------------------------------------------------------
# the batched C++ writer / reader
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["512x512_s0", "512x512_s2", "512x512_s4", "768x512_s0", "601x523_s0", "601x523_s2",
                                 "1000x700_s0", "530x300_s1"])
def test_choh_batch_single_file_is_stock_choh(tmp_path, key):
    dims, mode = key.split("_s")
    w, h = (int(v) for v in dims.split("x"))
    size, md5 = _golden(key)
    (tmp_path / "in.rgb").write_bytes(ol.synth_rgb(w, h, 1).tobytes())
    r = _run([_exe("choh_batch"), str(tmp_path / "in.rgb"), str(tmp_path / "out.hoh"), str(w), str(h), f"-s{mode}"])
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "out.hoh").read_bytes()
    assert r.stdout.strip() == str(size)
    assert len(data) == size and hashlib.md5(data).hexdigest() == md5


def test_choh_batch_many_images_equal_single_runs(tmp_path):
    """--batch: five 512x512 images in one GPU call; image 0 is the golden one, the others must equal what the
    per-file command line writes."""
    w = h = 512
    names = []
    for i in range(5):
        p = tmp_path / f"img{i}.rgb"
        p.write_bytes(ol.synth_rgb(w, h, 1 + i).tobytes())
        names.append(str(p))
    out = tmp_path / "out"
    out.mkdir()
    r = _run([_exe("choh_batch"), "--batch", str(w), str(h), "-s0", str(out)] + names)
    assert r.returncode == 0, r.stderr
    size, md5 = _golden("512x512_s0")
    first = (out / "img0.hoh").read_bytes()
    assert len(first) == size and hashlib.md5(first).hexdigest() == md5
    sizes = [int(v) for v in r.stdout.split()]
    for i in range(1, 5):
        single = tmp_path / f"single{i}.hoh"
        r1 = _run([_exe("choh_batch"), names[i], str(single), str(w), str(h), "-s0"])
        assert r1.returncode == 0
        assert single.read_bytes() == (out / f"img{i}.hoh").read_bytes()
        assert sizes[i] == len(single.read_bytes())


@pytest.mark.parametrize("w,h,mode", [(512, 512, 0), (601, 523, 1), (530, 300, 2), (512, 256, 3), (96, 80, 4), (256, 256, 2)])
def test_batch_writer_reader_round_trip(tmp_path, w, h, mode):
    """choh_batch --decodable -> dhoh_batch: exact RGB back, tiled and untiled images, images with LZ matches."""
    rng = np.random.default_rng(7 * mode + w)
    names, imgs = [], []
    for i in range(3):
        img = ol.photo_with_repeats(rng, w, h, 900 + 13 * mode + i)
        p = tmp_path / f"p{i}.rgb"
        p.write_bytes(img.tobytes())
        names.append(str(p))
        imgs.append(img.tobytes())
    enc = tmp_path / "enc"
    dec = tmp_path / "dec"
    enc.mkdir()
    dec.mkdir()
    r = _run([_exe("choh_batch"), "--batch", "--decodable", str(w), str(h), f"-s{mode}", str(enc)] + names)
    assert r.returncode == 0, r.stderr
    r = _run([_exe("dhoh_batch"), "--batch", str(dec)] + [str(enc / f"p{i}.hoh") for i in range(3)])
    assert r.returncode == 0, r.stderr
    for i in range(3):
        assert (dec / f"p{i}.rgb").read_bytes() == imgs[i], i
        assert len((enc / f"p{i}.hoh").read_bytes()) < len(imgs[i])


def test_dhoh_batch_decodes_a_stock_choh_file(tmp_path):
    """The file the real `choh -s0` wrote (tests/golden) -> the image it was made from: something the reference's
    own dhoh cannot do."""
    z = np.load(os.path.join(G, "layer_tile.npz"))
    (tmp_path / "stock.hoh").write_bytes(z["file_512x512_s0"].tobytes())
    r = _run([_exe("dhoh_batch"), str(tmp_path / "stock.hoh"), str(tmp_path / "back.rgb")])
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "back.rgb").read_bytes() == ol.synth_rgb(512, 512, 1).tobytes()


def test_dhoh_batch_reports_damage(tmp_path):
    z = np.load(os.path.join(G, "layer_tile.npz"))
    data = bytearray(z["file_512x512_s0"].tobytes())
    data[200000] ^= 0x40
    (tmp_path / "bad.hoh").write_bytes(bytes(data))
    r = _run([_exe("dhoh_batch"), str(tmp_path / "bad.hoh"), str(tmp_path / "bad.rgb")])
    assert r.returncode == 4 and "cannot be decoded" in r.stderr
    (tmp_path / "hdr.hoh").write_bytes(bytes([0x99, 0x48, 0x4f, 0x48, 0x02, 0x08, 0x01, 0x01]))
    r = _run([_exe("dhoh_batch"), str(tmp_path / "hdr.hoh"), str(tmp_path / "x.rgb")])
    assert r.returncode == 3 and "D1" in r.stderr
