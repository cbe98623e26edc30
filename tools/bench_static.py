#!/usr/bin/env python
"""BASELINE config 4: rans64 symbol-stream sweep over an 8-bit alphabet with ONE static frequency table
(prob_bits 12, built by normalize_freqs from the global histogram), symbols cut into 65 536-symbol streams,
2^20 ... 2^32 symbols in powers of 4.

    python tools/bench_static.py [--max-log2 32] [--cpu-streams 64]

GPU: hoh_rans_encode_static / hoh_rans_decode_static on device-resident symbols (u16, as the reference's
encoder takes them), round trip checked.  CPU: the reference's Rans64 loops (oracle/_ref) on the host cores,
one process per core, on a sample of streams.  One JSON line per size.  Development tool."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

STREAM, PB = 65536, 12


def _cpu(args):
    seed, count = args
    import oracle_lib as ol
    sym8 = ol.synth_symbols(STREAM * count, seed)
    f = np.bincount(sym8, minlength=256).astype(np.uint32)
    cum = np.zeros(257, np.uint32)
    ol.oracle().orc_normalize_freqs(f, cum, 256, 1 << PB)
    sym = sym8.astype(np.uint16)
    L = ol.ref() if ol.have_ref() else ol.oracle()
    enc = L.ref_rans_encode_static if ol.have_ref() else L.orc_rans_encode_static
    dec = L.ref_rans_decode_static if ol.have_ref() else L.orc_rans_decode_static
    buf = np.zeros(STREAM * 2 + 64, np.uint8)
    out = np.zeros(STREAM, np.uint16)
    t_enc = t_dec = 0.0
    for i in range(count):
        part = sym[i * STREAM:(i + 1) * STREAM]
        t0 = time.perf_counter()
        n = enc(part, STREAM, f, cum, 256, PB, buf)
        t1 = time.perf_counter()
        dec(buf, n, STREAM, f, cum, 256, PB, out)
        t2 = time.perf_counter()
        t_enc += t1 - t0
        t_dec += t2 - t1
    return t_enc, t_dec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=32)
    ap.add_argument("--cpu-streams", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import oracle_lib as ol
    mod = bench._load("hohgpu", os.path.join(ROOT, "hoh-ans_b200", "host", "hohgpu.py"))
    g = mod.HohGpu(0)
    cores = os.cpu_count() or 1
    per = max(1, a.cpu_streams // cores)
    with ProcessPoolExecutor(cores) as ex:
        list(ex.map(_cpu, [(1, 1)] * cores))
        res = list(ex.map(_cpu, [(3 + i, per) for i in range(cores)]))
    cpu_enc = per * STREAM * cores / max(r[0] for r in res) / 1e6
    cpu_dec = per * STREAM * cores / max(r[1] for r in res) / 1e6
    base = ol.synth_symbols(1 << 24, 7)  # 16 Mi symbols of the geometric source, tiled to the wanted size on the device
    f = np.bincount(base, minlength=256).astype(np.uint32)
    cum = np.zeros(257, np.uint32)
    assert ol.oracle().orc_normalize_freqs(f, cum, 256, 1 << PB) == 0
    d_cum = g.alloc(cum.nbytes).upload(cum)
    slab = (STREAM * PB // 8 + 64 + 15) & ~15
    for lg in range(20, a.max_log2 + 1, 2):
        n = 1 << lg
        n_streams = n // STREAM
        sym = base.astype(np.uint16)
        d_sym = g.alloc(n * 2 + 64)
        reps = max(1, n // sym.size)
        piece = sym[:min(n, sym.size)]
        for r in range(reps):
            g._ck(g.lib.hoh_h2d(g.ctx, d_sym.ptr + r * piece.nbytes, piece.ctypes.data, piece.nbytes), "h2d")
        d_out = g.alloc(n_streams * slab)
        d_len = g.alloc(n_streams * 4)
        d_dec = g.alloc(n * 2 + 64)

        def enc():
            g._ck(g.lib.hoh_rans_encode_static(g.ctx, d_sym.ptr, n, STREAM, d_cum.ptr, 256, PB, d_out.ptr, slab, d_len.ptr), "enc")

        def dec():
            g._ck(g.lib.hoh_rans_decode_static(g.ctx, d_out.ptr, slab, d_len.ptr, n, STREAM, d_cum.ptr, 256, PB, d_dec.ptr), "dec")
        enc(); dec(); g.sync()
        g.timer_start(0)
        for _ in range(a.steps):
            enc()
        g.timer_stop(0)
        g.timer_start(1)
        for _ in range(a.steps):
            dec()
        g.timer_stop(1)
        e_ms, d_ms = g.timer_ms(0) / a.steps, g.timer_ms(1) / a.steps
        lens = d_len.download(np.uint32, n_streams)
        back = d_dec.download(np.uint16, min(n, 1 << 24))
        ok = bool(np.array_equal(back, sym[:back.size]))
        comp = int(lens.astype(np.uint64).sum())
        alg = n * 1 + comp  # 1 B per symbol + H/8 B
        print(json.dumps({"config": "rans64 static-table sweep", "symbols_log2": lg, "streams": n_streams,
                          "encode_ms": e_ms, "decode_ms": d_ms, "encode_msym_s": n / e_ms / 1e3, "decode_msym_s": n / d_ms / 1e3,
                          "bits_per_symbol": comp * 8 / n, "roundtrip_ok": ok,
                          "algorithmic_gbs_encode": alg / e_ms / 1e6, "algorithmic_gbs_decode": alg / d_ms / 1e6,
                          "cpu_reference_msym_s": {"encode": cpu_enc, "decode": cpu_dec, "cores": cores,
                                                   "sample_streams": per * cores}}))
        for b in (d_sym, d_out, d_len, d_dec):
            b.free()


if __name__ == "__main__":
    main()
