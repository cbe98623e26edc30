#!/usr/bin/env python
"""Round trips at scales and shapes the test suite does not reach (one B200, a few minutes): one very large image,
many tiny images, a batch larger than one device chunk through the host-buffer calls, odd sizes at every mode.
Prints one line per case; exits 1 on the first mismatch.  usage: python tools/stress_shapes.py [--quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (fill_images: the generator of SURVEY 8(d), C, multi-threaded)
import gpu_lib  # noqa: E402

FIX_ENCODER = 24


def images(n, w, h, seed=1):
    rgb = np.zeros(n * w * h * 3, np.uint8)
    bench.fill_images(rgb, seed, n, w, h, os.cpu_count() or 1)
    return rgb


def case_s0(g, n, w, h):
    rgb = images(n, w, h)
    t0 = time.perf_counter()
    packed, off, res = g.encode_images_s0(rgb, n, w, h)
    assert (res["status"] == 0).all(), "encode status"
    back, st = g.decode_images_s0(packed, off, n, w, h)
    ok = bool((st == 0).all() and np.array_equal(back, rgb))
    print(f"s0      n={n:<7} {w}x{h:<6} ratio {len(packed) / rgb.size:.3f}  {time.perf_counter() - t0:6.1f} s  {'ok' if ok else 'MISMATCH'}", flush=True)
    return ok


def case_host(g, n, w, h, mode):
    rgb = images(n, w, h)
    t0 = time.perf_counter()
    packed, off, rec = g.encode_images_host(rgb, n, w, h, mode, FIX_ENCODER)
    assert (rec["status"] == 0).all(), "encode status"
    back, st = g.decode_images_host(packed, off, n, w, h)
    ok = bool((st == 0).all() and np.array_equal(np.asarray(back), rgb))
    print(f"host m{mode} n={n:<7} {w}x{h:<6} ratio {len(packed) / rgb.size:.3f}  {time.perf_counter() - t0:6.1f} s  {'ok' if ok else 'MISMATCH'}", flush=True)
    g.host_free(back)
    return ok


def guarded(fn):
    def run(*a):
        try:
            return fn(*a)
        except Exception as e:  # report and go on: the point is to find every shape that breaks
            print(f"{fn.__name__}{a[1:]}: {type(e).__name__}: {e}", flush=True)
            return False
    return run


def main():
    global case_s0, case_host
    case_s0, case_host = guarded(case_s0), guarded(case_host)
    quick = "--quick" in sys.argv
    g = gpu_lib.gpu()
    ok = True
    # one very large image: 4096 tiles of 256x256 in ONE image (tile index arithmetic, 64-bit offsets inside an image)
    ok &= case_s0(g, 1, 16384, 16384 if not quick else 4096)
    # many tiny images: one 2x2 / 8x8 / 40x24 tile each (stored-mode streams, short chains, ragged warps)
    ok &= case_s0(g, 20000 if not quick else 2000, 2, 2)
    ok &= case_s0(g, 20000 if not quick else 2000, 8, 8)
    ok &= case_host(g, 5000 if not quick else 500, 40, 24, 2)
    # widths that are not multiples of 4 / 8 (generic residual kernel, unfused decoder, half-warp un-prediction)
    ok &= case_s0(g, 64, 1001, 701)
    ok &= case_host(g, 16, 1001, 701, 1)
    ok &= case_host(g, 8, 601, 523, 4)
    # a batch larger than one device chunk through the host-buffer calls at mode 0 (12.9 GB of pixels)
    if not quick:
        ok &= case_host(g, 16384, 512, 512, 0)
    g.close()
    print("all ok" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
