"""Contexts on two GPUs in ONE process: every entry point runs on its context's device whatever the caller's current
device is, and leaves the caller's device as it found it (DeviceGuard in hoh_api.cu).  Skipped on a one-GPU box
(the driver's per-rank bench never needs it: one process per GPU)."""
import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_two_contexts_two_devices_interleaved():
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    mod = gpu_lib.hohgpu()
    g0, g1 = mod.HohGpu(0), mod.HohGpu(1)
    try:
        w = h = 512
        n = 3
        a = np.concatenate([ol.synth_rgb(w, h, 11 + i) for i in range(n)])
        b = np.concatenate([ol.synth_rgb(w, h, 51 + i) for i in range(n)])
        torch.cuda.set_device(0)
        pa, oa, ra = g0.encode_images_s0(a, n, w, h)
        pb, ob, rb = g1.encode_images_s0(b, n, w, h)  # the caller's current device is 0
        assert torch.cuda.current_device() == 0
        torch.cuda.set_device(1)
        back_a, sa = g0.decode_images_s0(pa, oa, n, w, h)  # ... and now 1
        assert torch.cuda.current_device() == 1
        back_b, sb = g1.decode_images_s0(pb, ob, n, w, h)
        assert (ra["status"] == 0).all() and (rb["status"] == 0).all()
        assert (sa == 0).all() and (sb == 0).all()
        assert np.array_equal(back_a, a) and np.array_equal(back_b, b)
        # the whole-tile codec at mode 2 (child contexts, side streams, scratch on the right device)
        small = np.concatenate([ol.synth_rgb(96, 80, 5 + i) for i in range(2)])
        t1, r1 = g1.encode_images(small, 2, 96, 80, 2, 24)
        t0, r0 = g0.encode_images(small, 2, 96, 80, 2, 24)
        assert t0 == t1
        d1, s1 = g1.decode_images(t1, 2, 96, 80)
        d0, s0 = g0.decode_images(t0, 2, 96, 80)
        assert (s0 == 0).all() and (s1 == 0).all()
        assert np.array_equal(d0, small) and np.array_equal(d1, small)
        assert torch.cuda.current_device() == 1
    finally:
        torch.cuda.set_device(0)
        g0.close()
        g1.close()
