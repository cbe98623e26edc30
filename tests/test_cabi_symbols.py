"""CPU-side checks of the drop-in boundary: libhohgpu.so builds for sm_100a, loads, exports every
symbol include/hohgpu.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import shutil

import pytest

import gpu_lib

ROOT = gpu_lib.ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "hohgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hoh_[a-z0-9_]+)\s*\(", text)))


def _lib_path():
    b = gpu_lib.builder()
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build()
    if not os.path.exists(b.LIB):
        pytest.skip("libhohgpu.so not built and nvcc absent")
    return b.LIB


def test_header_declares_the_documented_surface():
    names = _declared()
    for must in ("hoh_encode_entropy", "hoh_decode_entropy", "hoh_subtract_green", "hoh_channelpredict_fastpath",
                 "hoh_channelpredict_section", "hoh_channelpredict_all", "hoh_unpredict_all", "hoh_normalize_freqs",
                 "hoh_encode_images_s0", "hoh_decode_images_s0", "hoh_encode_entropy_batch",
                 "hoh_decode_entropy_batch", "hoh_rans_encode_static", "hoh_rans_decode_static"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_mirror_covers_every_symbol():
    mod = gpu_lib.hohgpu()
    assert sorted(mod.SIGNATURES) == _declared()


def test_pure_helpers_need_no_gpu():
    _lib_path()
    mod = gpu_lib.hohgpu()
    lib = mod.load_library()
    g = mod.TileGeometry()
    assert lib.hoh_tile_geometry_for(512, 512, C.byref(g)) == 0
    assert (g.x_tiles, g.y_tiles, g.tile_w, g.tile_h, g.streams_per_image) == (2, 2, 256, 256, 12)
    assert lib.hoh_tile_geometry_for(3840, 2160, C.byref(g)) == 0  # choh.cpp:455-460
    assert (g.x_tiles, g.y_tiles, g.tile_w, g.tile_h) == (15, 8, 256, 270)
    assert lib.hoh_tile_geometry_for(256, 256, C.byref(g)) == 0    # not tiled (choh.cpp:454)
    assert (g.x_tiles, g.y_tiles, g.tile_w, g.tile_h) == (1, 1, 256, 256)
    assert lib.hoh_tile_geometry_for(1000, 300, C.byref(g)) == 0
    assert (g.x_tiles, g.y_tiles, g.tile_w, g.tile_h) == (3, 1, 334, 300)
    assert lib.hoh_enc_slab_bytes(65536, 15) % 16 == 0
    assert lib.hoh_enc_slab_bytes(65536, 15) >= 65536 * 15 // 8 + 1536


def test_no_cpu_fallback_without_a_device():
    _lib_path()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    mod = gpu_lib.hohgpu()
    with pytest.raises(mod.HohError) as e:
        mod.HohGpu(0)
    assert e.value.status == mod.HOH_E_CUDA


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C-ABI: include/hohgpu.h compiles as C99 (-pedantic, no C++), and a C program that takes the
    address of entry points links against libhohgpu.so and runs (hoh_ctx_create fails cleanly without a device)."""
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc absent")
    import subprocess
    lib = _lib_path()
    src = tmp_path / "c_user.c"
    src.write_text(
        '#include <stdio.h>\n#include "hohgpu.h"\n'
        "typedef void (*fn)(void);\n"
        "static fn volatile table[] = {(fn)hoh_encode_images_host, (fn)hoh_decode_images_host, (fn)hoh_encode_images_s0};\n"
        "int main(void) {\n"
        "    hoh_ctx* ctx = NULL;\n"
        "    int st = hoh_ctx_create(0, NULL, &ctx);\n"
        '    printf("%d %s\\n", st, hoh_strerror(st));\n'
        "    if (st == HOH_OK) hoh_ctx_destroy(ctx);\n"
        "    return table[0] == 0 || table[1] == 0 || table[2] == 0;\n"
        "}\n")
    exe = tmp_path / "c_user"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                           "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.dirname(lib), "-lhohgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    st = int(out.stdout.split()[0])
    assert st in (0, 1), out.stdout  # HOH_OK on a GPU box, HOH_E_CUDA here: never a crash, never a CPU fallback
