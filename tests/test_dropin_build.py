"""The drop-in boundary builds (CPU only, no compute): include/hoh_gpu_shim.hpp force-included in front of the
reference's OWN, unedited host programs compiles and links against libhohgpu.so; the repo's batched C++ container
writer / reader build; and every one of those programs refuses to run without a GPU instead of falling back.
The command line INTEGRATION.md gives a maintainer is read out of that file and run verbatim."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
LIB = os.path.join(ROOT, "hoh-ans_b200", "csrc", "libhohgpu.so")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "choh.cpp")), reason="reference sources not present")


def _ensure_lib():
    import gpu_lib
    gpu_lib.builder().build()
    assert os.path.exists(LIB)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


INTEGRATION_CMD = ("g++ -O3 -include include/hoh_gpu_shim.hpp -Iinclude -I/root/reference /root/reference/choh.cpp "
                   "-o dropin/_bin/choh_gpu -Lhoh-ans_b200/csrc -lhohgpu -Wl,-rpath,'$ORIGIN/../../hoh-ans_b200/csrc' -lm -lrt")


@needs_ref
def test_integration_md_command_builds_choh_gpu():
    """the documented command, verbatim, run from the repo root."""
    _ensure_lib()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    cmds = [" ".join(m.split()) for m in re.findall(r"```sh\n(g\+\+ [^`]*?choh\.cpp[^`]*?)\n```", text)]
    assert INTEGRATION_CMD in cmds, cmds
    os.makedirs(os.path.join(ROOT, "dropin", "_bin"), exist_ok=True)
    subprocess.run(INTEGRATION_CMD, shell=True, cwd=ROOT, check=True, capture_output=True)
    assert os.access(os.path.join(ROOT, "dropin", "_bin", "choh_gpu"), os.X_OK)


@needs_ref
def test_every_reference_program_relinks_against_the_gpu(tmp_path):
    """choh, dhoh, simple_entropy_encoder / _decoder and layer_roundtrip_test from the reference tree, unedited."""
    _ensure_lib()
    subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "-B", "ref", f"BIN={tmp_path}"], check=True,
                   capture_output=True)
    built = sorted(os.listdir(tmp_path))
    assert built == ["choh_gpu", "dhoh_gpu", "layer_roundtrip_gpu", "simple_entropy_decoder_gpu",
                     "simple_entropy_encoder_gpu"]
    # the relinked tools carry no copy of the hot path: the reference's own symbols for it are gone ...
    syms = subprocess.run(["nm", "-C", str(tmp_path / "choh_gpu")], capture_output=True, text=True, check=True).stdout
    assert "Rans64EncPutSymbol" not in syms and "hoh_encode_entropy" in syms and "hoh_channelpredict_section" in syms
    # ... and they load libhohgpu.so at run time
    ldd = subprocess.run(["ldd", str(tmp_path / "choh_gpu")], capture_output=True, text=True).stdout
    assert "libhohgpu.so" in ldd
    subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "ref"], check=True, capture_output=True)  # the in-tree copies ship to the GPU box


def test_batch_tools_build(tmp_path):
    _ensure_lib()
    subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "-B", "tools", f"BIN={tmp_path}"], check=True,
                   capture_output=True)
    assert sorted(os.listdir(tmp_path)) == ["choh_batch", "dhoh_batch"]
    subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "tools"], check=True, capture_output=True)


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a machine without a GPU")
def test_programs_fail_loudly_without_a_gpu(tmp_path):
    """No CPU fallback anywhere behind the boundary: exit code 3 and a message, never an output file."""
    _ensure_lib()
    subprocess.run(["make", "-C", os.path.join(ROOT, "dropin"), "all"], check=True, capture_output=True)
    rgb = tmp_path / "in.rgb"
    rgb.write_bytes(bytes(range(48)))
    out = tmp_path / "out.hoh"
    for exe in ("choh_batch", "choh_gpu"):
        path = os.path.join(ROOT, "dropin", "_bin", exe)
        if not os.path.exists(path):
            continue
        r = subprocess.run([path, str(rgb), str(out), "4", "4", "-s0"], capture_output=True, text=True)
        assert r.returncode == 3, (exe, r.returncode, r.stderr)
        assert "no CPU fallback" in r.stderr
        assert not out.exists()
    hoh = tmp_path / "in.hoh"
    hoh.write_bytes(bytes([153, 72, 79, 72, 2, 8, 3, 3, 0, 0, 128, 3]))
    r = subprocess.run([os.path.join(ROOT, "dropin", "_bin", "dhoh_batch"), str(hoh), str(tmp_path / "o.rgb")],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
