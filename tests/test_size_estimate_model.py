"""The payload length of a Rans64 stream from its histogram and table alone (warp_estimate_words in
hoh-ans_b200/csrc/hoh_kernels.cuh): the bound restated in Python and checked against a step-by-step model of the coder
(rans64.hpp:77-94) on residual planes and on skewed synthetic streams, prob_bits 12..19.  The GPU path checks every stream
it codes against its interval as well (HOH_S_BAD_ESTIMATE in k_finish_streams)."""
import math

import numpy as np

import oracle_lib as ol


def _normalised(hist, bits):
    f = hist.astype(np.uint32).copy()
    cum = np.zeros(len(f) + 1, np.uint32)
    assert ol.oracle().orc_normalize_freqs(f, cum, len(f), 1 << bits) == 0
    return f, cum


def _coder_words(symbols, f, cum, bits):
    """Words emitted by Rans64EncPut over the stream, last symbol first (entropy_encoding.hpp:222-225)."""
    x, words = 1 << 31, 0
    fl, cl = f.tolist(), cum.tolist()
    for s in symbols[::-1].tolist():
        fr = fl[s]
        if x >= ((1 << (31 - bits)) << 32) * fr:  # rans64.hpp:82-86
            words += 1
            x >>= 32
        x = ((x // fr) << bits) + (x % fr) + cl[s]
    assert (1 << 31) <= x < (1 << 63)
    return words


def _estimate(hist, f, bits):
    """warp_estimate_words, same arithmetic."""
    S = B = 0.0
    for c, fr in zip(hist.tolist(), f.tolist()):
        if c:
            S += c * (bits - math.log2(fr))
            B += c * fr
    B *= 1.45 / 2147483648.0
    k = 1.45 * 2.0 ** (bits - 31)
    w_cap = math.floor((S + B + 1.0) / 32.0) + 1.0
    slack = 1e-4 + 1e-12 * S
    e_hi, e_lo = B + w_cap * k + slack, -(B + 2.0 * w_cap * k + slack)
    return max(0, math.floor((S + e_lo) / 32.0)), math.floor((S + e_hi) / 32.0)


def _streams():
    rng = np.random.default_rng(5)
    out = []
    for seed in (1, 2):  # the residual planes of a synthetic photograph (G at 8 bits, R-G at 9)
        rgb = ol.synth_rgb(128, 96, seed)
        planes = [np.zeros(128 * 96, np.uint16) for _ in range(3)]
        ol.oracle().orc_subtract_green(rgb, rgb.size, *planes)
        out.append((planes[0], 256))
        out.append((planes[1], 512))
    for k in range(4):  # skewed alphabets, from two dominant symbols to nearly flat
        p = rng.dirichlet(np.full(48, 0.03 + 0.3 * k))
        out.append(((rng.choice(48, 20000, p=p) + 200).astype(np.uint16), 512))
    out.append((np.full(5000, 7, np.uint16) + (rng.random(5000) < 0.001), 256))  # almost one symbol
    return out


def test_interval_contains_the_coded_length():
    total = open_ = 0
    for symbols, range_ in _streams():
        hist = np.bincount(symbols, minlength=range_)
        for bits in (12, 14, 15, 16, 17, 19):
            f, cum = _normalised(hist, bits)
            words = _coder_words(symbols, f, cum, bits)
            lo, hi = _estimate(hist, f, bits)
            assert lo <= words <= hi, (range_, bits, words, lo, hi)
            assert hi - lo <= 1
            total += 1
            open_ += hi != lo
    assert open_ * 3 <= total  # most lengths are known exactly without coding
