"""Two host threads, one context each, on the same GPU at the same time (hohgpu.h: entry points are thread-safe per
context handle; ctypes releases the GIL during the calls, so the library really runs concurrently)."""
import threading

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_two_threads_two_contexts_one_gpu():
    mod = gpu_lib.hohgpu()
    errors = []

    def work(k):
        g = mod.HohGpu(0)
        try:
            w = h = 512
            n = 4
            rgb = np.concatenate([ol.synth_rgb(w, h, 100 * k + i) for i in range(n)])
            small = np.concatenate([ol.synth_rgb(96, 80, 7 * k + i) for i in range(3)])
            want_small = None
            for _ in range(3):
                packed, off, res = g.encode_images_s0(rgb, n, w, h)
                back, st = g.decode_images_s0(packed, off, n, w, h)
                assert (res["status"] == 0).all() and (st == 0).all() and np.array_equal(back, rgb)
                tiles, rec = g.encode_images(small, 3, 96, 80, 2 + k, 24)
                want_small = want_small or tiles
                assert tiles == want_small  # the same bytes every time, whatever the other thread is doing
                d, s = g.decode_images(tiles, 3, 96, 80)
                assert (s == 0).all() and np.array_equal(d, small)
        except BaseException as e:  # noqa: BLE001 - reported in the main thread
            errors.append((k, repr(e)))
        finally:
            g.close()

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
