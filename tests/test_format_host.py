"""Pins hoh-ans_b200/csrc/hoh_format.cuh (the stream-format code a GPU lane runs per stream) against
the oracle, on the CPU: the header is compiled for the host by tests/hostfmt/fmt_host.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostfmt", "fmt_host.cpp")
SO = os.path.join(HERE, "hostfmt", "libfmt_host.so")
HDR = os.path.join(os.path.dirname(HERE), "hoh-ans_b200", "csrc", "hoh_format.cuh")


def _lib():
    if (not os.path.exists(SO)) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-frounding-math", "-fPIC", "-shared", "-x", "c++", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.fmt_build_head.restype = C.c_uint32
    L.fmt_build_head.argtypes = [ol.u32p, C.c_uint32, C.c_uint32, C.c_uint32, ol.u8p, C.POINTER(C.c_uint32)]
    L.fmt_parse.restype = C.c_uint64
    L.fmt_parse.argtypes = [ol.u8p, C.c_uint64, C.c_uint, ol.u32p, ol.u32p]
    L.fmt_table_mode.restype = C.c_uint32
    L.fmt_table_mode.argtypes = [ol.u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.fmt_div_model.restype = C.c_uint64
    L.fmt_div_model.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_int]
    L.fmt_put_varint.restype = C.c_uint32
    L.fmt_put_varint.argtypes = [ol.u8p, C.c_uint32, C.c_uint32]
    return L


def _norm(sym, rangev, pb):
    f = np.bincount(sym, minlength=rangev).astype(np.uint32)
    cum = np.zeros(rangev + 1, np.uint32)
    st = ol.oracle().orc_normalize_freqs(f, cum, rangev, 1 << pb)
    return f, cum, st


def test_varint_matches_oracle():
    L = _lib()
    for v in [0, 1, 127, 128, 300, 16383, 16384, 65536, (1 << 21) - 1, 1 << 21, 5000000]:
        a = np.zeros(8, np.uint8)
        b = np.zeros(8, np.uint8)
        na = L.fmt_put_varint(a, 0, v)
        nb = ol.oracle().orc_write_varint(b, 0, v)
        assert na == nb and a.tobytes() == b.tobytes(), v


def test_head_and_parse_match_oracle_streams():
    L = _lib()
    rng = np.random.default_rng(5)
    checked = 0
    for it in range(1200):
        rangev = int(rng.choice([1, 2, 3, 5, 14, 16, 255, 256, 257, 512]))
        pb = int(rng.integers(8, 20))
        if (1 << pb) < rangev:
            pb = 10
        n = int(rng.choice([1, 2, 7, 20, 49, 100, 1000, 5000, 70000]))
        kind = it % 6
        if kind == 0:
            sym = rng.integers(0, rangev, n)
        elif kind == 1:
            sym = np.clip(np.rint(rng.laplace(rangev / 2, rangev / 40 + 0.3, n)), 0, rangev - 1)
        elif kind == 2:
            sym = np.full(n, int(rng.integers(0, rangev)))
        elif kind == 3:
            sym = np.clip(rng.geometric(0.3, n) - 1, 0, rangev - 1)
        elif kind == 4:
            sym = np.clip(np.rint(rng.laplace(rangev / 2, 1.0, n)), 0, rangev - 1)
        else:
            sym = rng.integers(0, max(1, rangev // 8), n)
        sym = sym.astype(np.uint16)
        f, cum, st = _norm(sym, rangev, pb)
        if st != 0:
            continue
        want, st2 = ol.orc_encode_entropy(sym, rangev, pb)
        assert st2 == 0
        head = np.zeros(4096, np.uint8)
        stored = C.c_uint32(0)
        hl = L.fmt_build_head(f, rangev, n, pb, head, C.byref(stored))
        r, nn, em, pb5, tm = ol.peek_stream(want)
        if em == 1:  # rANS form: our head must be the exact prefix, followed by varint(payload)
            assert want[:hl].tobytes() == head[:hl].tobytes(), (it, rangev, pb, n)
            assert stored.value >= len(want)
            if f.max() >= (1 << pb):  # a symbol owning all of 2^pb overflows its field and the
                continue              # carry corrupts the clamps: unparseable in the reference too
            # parse it back
            fields = np.zeros(8, np.uint32)
            pf = np.zeros(rangev, np.uint32)
            padded = np.concatenate([want, np.zeros(16, np.uint8)])
            end = L.fmt_parse(padded, 0, 7, fields, pf)
            assert end == hl
            assert (fields[0], fields[1], fields[2], fields[3], fields[4]) == (rangev, n, 1, pb, tm)
            # table mode 1 is lossy when a frequency needs more than maxbits bits (D6): compare
            # against what the oracle's own decoder reads by decoding the symbols instead
            if tm == 2:
                assert np.array_equal(pf, f), (it, rangev, pb, n)
            checked += 1
        else:
            assert stored.value == len(want)
    assert checked > 200


def test_representable_rule_only_changes_lossy_mode1_tables():
    """HOH_FIX_LONE's table rule (plan_head's `representable`): table mode 1 writes every frequency on
    bit_length(range-1) bits, so it is replaced by mode 2 exactly when some frequency does not fit;
    every other decision is the reference's."""
    L = _lib()
    rng = np.random.default_rng(12)
    changed = kept1 = 0
    for it in range(3000):
        rangev = int(rng.choice([2, 3, 5, 9, 14, 16, 40, 256]))
        pb = int(rng.integers(max(4, int(np.ceil(np.log2(rangev)))), 13))
        n = int(rng.integers(rangev, 400))
        sym = np.minimum(rng.geometric(float(rng.uniform(0.05, 0.9)), n) - 1, rangev - 1).astype(np.uint16)
        f, cum, st = _norm(sym, rangev, pb)
        if st:
            continue
        m0 = L.fmt_table_mode(f, rangev, n, pb, 0)
        m1 = L.fmt_table_mode(f, rangev, n, pb, 1)
        maxbits = int(rangev - 1).bit_length()
        lossy = m0 == 1 and bool((f >> maxbits).any())
        assert m1 == (2 if lossy else m0), (it, rangev, pb, m0, m1)
        changed += lossy
        kept1 += (m0 == 1 and not lossy)
    assert changed > 20 and kept1 > 20


def test_encoder_division_step_model_is_exact():
    """The rANS encoder's x / freq for prob_bits >= 14 is one fused multiply-add rounded towards zero whose
    mantissa is the quotient (hoh_kernels.cuh, rans_put<0> / <2>): its CPU model against the integer division, over
    random states, the top of the state range and both ends of quotient intervals, for reciprocals up to two
    ulps either side of the nominal one."""
    L = _lib()
    for bits in range(12, 20):  # 12-13: rans_put<2>, up to three fix-ups; 14-19: rans_put<0>, one
        for ulps in (-2, 0, 2):
            assert L.fmt_div_model(bits, 400000, 1234567 + bits, ulps) == 0, (bits, ulps)
