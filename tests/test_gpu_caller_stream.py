"""A context bound to the CALLER's CUDA stream (hoh_ctx_create(device, stream)): the library's work is ordered on that
stream — a torch tensor filled on it is encoded without any host synchronisation in between, and the decode's output is
read back through the same stream."""
import ctypes as C

import numpy as np
import pytest

import gpu_lib
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def test_context_on_a_torch_stream():
    torch = pytest.importorskip("torch")
    mod = gpu_lib.hohgpu()
    stream = torch.cuda.Stream(device=0)
    g = mod.HohGpu(0, C.c_void_p(stream.cuda_stream))
    try:
        w = h = 512
        n = 2
        rgb = np.concatenate([ol.synth_rgb(w, h, 40 + i) for i in range(n)])
        geo = g.tile_geometry(w, h)
        n_streams = n * geo.streams_per_image
        out_bytes = int(g.lib.hoh_encode_images_out_bytes(C.byref(geo), n))
        packed_cap = rgb.size * 2 + 4096 * n_streams
        host = torch.from_numpy(rgb).pin_memory()
        with torch.cuda.stream(stream):
            d_rgb = torch.empty(rgb.size, dtype=torch.uint8, device="cuda:0")
            d_rgb.copy_(host, non_blocking=True)           # queued on the caller's stream, not waited for
            d_out = torch.empty(out_bytes, dtype=torch.uint8, device="cuda:0")
            d_res = torch.empty(n_streams * mod.RESULT_DT.itemsize, dtype=torch.uint8, device="cuda:0")
            d_packed = torch.zeros(packed_cap, dtype=torch.uint8, device="cuda:0")
            d_off = torch.empty(n_streams + 1, dtype=torch.int64, device="cuda:0")
            d_back = torch.zeros(rgb.size, dtype=torch.uint8, device="cuda:0")
            d_st = torch.empty(n_streams, dtype=torch.int32, device="cuda:0")
            g._ck(g.lib.hoh_encode_images_s0(g.ctx, d_rgb.data_ptr(), n, w, h, None, d_out.data_ptr(), out_bytes,
                                             d_res.data_ptr(), d_packed.data_ptr(), packed_cap, d_off.data_ptr()),
                  "hoh_encode_images_s0")
            g._ck(g.lib.hoh_decode_images_s0(g.ctx, d_packed.data_ptr(), packed_cap, d_off.data_ptr(), n, w, h, None,
                                             d_back.data_ptr(), d_st.data_ptr()), "hoh_decode_images_s0")
            back = d_back.cpu()                              # ordered after the decode on the same stream
            st = d_st.cpu()
            off = d_off.cpu()
        stream.synchronize()
        assert (st.numpy() == 0).all()
        assert np.array_equal(back.numpy(), rgb)
        assert 0 < int(off[-1]) < rgb.size
    finally:
        g.close()
