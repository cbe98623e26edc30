// TEST INFRASTRUCTURE: compiles the product's stream-format header (hoh-ans_b200/csrc/hoh_format.cuh)
// for the CPU so that the byte layout logic one GPU lane runs per stream can be pinned against the
// oracle without a GPU.  Not linked into libhohgpu.so.
#include "../../hoh-ans_b200/csrc/hoh_format.cuh"
#include <stddef.h>

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
}
