// TEST INFRASTRUCTURE: compiles the product's stream-format header (hoh-ans_b200/csrc/hoh_format.cuh)
// for the CPU so that the byte layout logic one GPU lane runs per stream can be pinned against the
// oracle without a GPU.  Not linked into libhohgpu.so.
#include "../../hoh-ans_b200/csrc/hoh_format.cuh"
#include <stddef.h>
#include <cfenv>
#include <cmath>
#include <cstring>

extern "C" {
uint32_t fmt_build_head(const uint32_t* freqs, uint32_t range, uint32_t n, uint32_t prob_bits,
                        uint8_t* head, uint32_t* stored_size) {
    return hohfmt::build_head(freqs, range, n, prob_bits, head, stored_size);
}
// returns offset after the table; fills header fields and freqs (modes 1/2 only)
uint64_t fmt_parse(const uint8_t* in, uint64_t at, unsigned flags, uint32_t* fields, uint32_t* freqs) {
    hohfmt::StreamHead h = hohfmt::parse_head(in, at, flags);
    fields[0] = h.range; fields[1] = h.n; fields[2] = h.rans; fields[3] = h.prob_bits;
    fields[4] = h.table_mode; fields[5] = h.empty; fields[6] = (uint32_t)h.body;
    if (h.empty || !h.rans || h.table_mode == 0 || h.table_mode == 3) return h.body;
    return hohfmt::parse_table(in, h, freqs);
}
uint32_t fmt_put_varint(uint8_t* dst, uint32_t at, uint32_t v) { return hohfmt::put_varint(dst, at, v); }
// table mode plan_head picks, with and without the "representable tables only" rule (HOH_FIX_LONE)
uint32_t fmt_table_mode(const uint32_t* freqs, uint32_t range, uint32_t n, uint32_t prob_bits, uint32_t representable) {
    uint8_t head[16];
    uint32_t stored, mode;
    hohfmt::ClampSet cs;
    hohfmt::plan_head(freqs, range, n, prob_bits, head, &stored, &cs, &mode, representable);
    return mode;
}
// CPU model of the encoder's division step for prob_bits >= 12 (hoh_kernels.cuh rans_put<0> / rans_put<2>):
// q = mantissa(fma_rz(double_rz(x), inv, 2^52)), r = lo(x) - lo(q) * f, one upward fix-up.  `inv` is modelled as
// RN(1/f) * (1 - 2^-50) moved by `ulps` (the device's Newton result may differ from RN(1/f) by an ulp or two).
// Walks states below x_max = f << (63 - bits) — random ones, the top of the range, both ends of quotient
// intervals, the post-renormalisation range — and returns how many disagree with the integer division.
uint64_t fmt_div_model(uint32_t bits, uint64_t cases, uint64_t seed, int ulps) {
    uint64_t s = seed ? seed : 88172645463325252ull, bad = 0;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (uint64_t it = 0; it < cases; it++) {
        uint32_t f = 1u + (uint32_t)(rnd() % (1u << bits));
        if ((it & 7) == 0) f = 1u + (uint32_t)(rnd() % 4);
        if ((it & 7) == 1) f = (1u << bits) - (uint32_t)(rnd() % 4);
        std::fesetround(FE_TONEAREST);
        volatile double one_over = 1.0 / (double)f;
        volatile double inv = one_over * 0.99999999999999911182;
        for (int u = 0; u < (ulps < 0 ? -ulps : ulps); u++) inv = std::nextafter((double)inv, ulps < 0 ? 0.0 : 1.0);
        const uint64_t xmax = (uint64_t)f << (63 - bits);
        uint64_t x;
        switch (it & 3) {
            case 0: x = rnd() % xmax; break;
            case 1: x = xmax - 1 - (rnd() % 1000) % xmax; break;
            case 2: {
                const uint64_t k = rnd() % (xmax / f ? xmax / f : 1);
                x = k * f + ((rnd() & 1) ? f - 1 : 0);
                if (x >= xmax) x = xmax - 1;
            } break;
            default: x = rnd() % (1ull << 31); break;
        }
        std::fesetround(FE_TOWARDZERO);
        volatile double xd = (double)x;
        volatile double t = std::fma((double)xd, (double)inv, 4503599627370496.0);
        const double tt = t;
        uint64_t b;
        std::memcpy(&b, &tt, 8);
        uint64_t q = b & 0x000fffffffffffffull;
        uint32_t r = (uint32_t)x - (uint32_t)q * f;
        if (r >= f) { r -= f; q++; }
        if (bits < 14)  // rans_put<2>: prob_bits 12-13, the estimate may be short by a few more; the device loops
            for (int k = 0; k < 6 && r >= f; k++) { r -= f; q++; }
        bad += !(q == x / f && r == x % f);
    }
    std::fesetround(FE_TONEAREST);
    return bad;
}
}
