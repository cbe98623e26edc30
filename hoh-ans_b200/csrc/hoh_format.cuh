// hoh_format.cuh — byte-exact stream format of hoh-ANS entropy streams, as straight-line code that
// one GPU lane runs per stream (header varints, metadata byte, frequency-table serialisation and
// parsing).  Everything here is __host__ __device__ so that tests/ can compile the very same source
// for the CPU and pin it against the oracle without a GPU; the product only ever calls it from
// kernels (hoh_kernels.cu).
//
// Reference behaviour restated here (file:line into the reference tree):
//   varint.hpp:6-45      read_varint / write_varint (7-bit groups, big-endian, at most 3 bytes)
//   varint.hpp:47-106    stuffer / unstuffer (MSB-first bit packing with ADD semantics)
//   entropy_encoding.hpp:24-27, 43-203   header, clamp search, table modes 1 and 2
//   entropy_decoding.hpp:143-253         header and table parse, modes 0 / 1 / 2
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HOH_HD __host__ __device__ __forceinline__
#else
#define HOH_HD inline
#endif

namespace hohfmt {

// ---- varints ---------------------------------------------------------------------------------
// varint.hpp:29-45.  Values >= 2^21 produce no bytes at all — kept.
HOH_HD uint32_t varint_len(uint32_t v) { return v < 128u ? 1u : (v < 16384u ? 2u : (v < 2097152u ? 3u : 0u)); }

HOH_HD uint32_t put_varint(uint8_t* dst, uint32_t at, uint32_t v) {
    uint32_t len = varint_len(v);
    for (uint32_t k = 0; k < len; k++) {
        uint32_t group = (v >> (7u * (len - 1u - k))) & 0x7fu;
        dst[at + k] = (uint8_t)(k + 1u < len ? (group | 0x80u) : group);
    }
    return at + len;
}

// varint.hpp:6-27.  The third byte is taken whole (its top bit is data).
template <typename Bytes>
HOH_HD uint32_t get_varint(const Bytes& src, uint64_t* at) {
    uint32_t b0 = src[(*at)++];
    if (!(b0 & 0x80u)) return b0;
    uint32_t b1 = src[(*at)++];
    if (!(b1 & 0x80u)) return ((b0 & 0x7fu) << 7) + b1;
    uint32_t b2 = src[(*at)++];
    return ((b0 & 0x7fu) << 14) + ((b1 & 0x7fu) << 7) + b2;
}

HOH_HD uint32_t bit_length(uint32_t v) {  // entropy_encoding.hpp:24-27
    uint32_t b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}

// ---- bit packer (varint.hpp:47-77) --------------------------------------------------------------
// The reference keeps a pending byte and ADDS each field into it without masking the field to its
// width, so an over-wide value carries into higher bits of the pending byte (mod 256).  `pending` /
// `room` play the roles of its remainder / bits_remaining.
struct BitSink {
    uint8_t* dst;
    uint32_t at;
    uint32_t pending;
    uint32_t room;  // free bits in pending, 1..8

    HOH_HD void open(uint8_t* d, uint32_t start) {
        dst = d;
        at = start;
        pending = 0;
        room = 8;
    }
    HOH_HD void small(uint32_t value, uint32_t bits) {  // bits <= 8
        if (bits < room) {
            pending = (pending + ((value << (room - bits)) & 0xffu)) & 0xffu;
            room -= bits;
        } else if (bits == room) {
            dst[at++] = (uint8_t)(pending + (value & 0xffu));
            pending = 0;
            room = 8;
        } else {
            uint32_t over = bits - room;
            dst[at++] = (uint8_t)(pending + ((value >> over) & 0xffu));
            room = 8 - over;
            pending = (value << room) & 0xffu;
        }
    }
    HOH_HD void put(uint32_t value, uint32_t bits) {
        // varint.hpp:65-70: a field wider than a byte goes out as its top (bits mod 8, or 8) bits
        // followed by whole low bytes.
        if (bits > 8) {
            uint32_t low_bytes = (bits - 1) / 8;
            small(value >> (8 * low_bytes), bits - 8 * low_bytes);
            while (low_bytes--) small((value >> (8 * low_bytes)) & 0xffu, 8);
        } else {
            small(value, bits);
        }
    }
    HOH_HD uint32_t close() {  // entropy_encoding.hpp:144-146
        if (room != 8) dst[at++] = (uint8_t)pending;
        return at;
    }
};

// ---- bit reader (varint.hpp:79-106) -------------------------------------------------------------
template <typename Bytes>
struct BitSource {
    Bytes src;
    uint64_t at;
    uint32_t held;
    uint32_t have;

    HOH_HD uint32_t get(uint32_t bits) {
        uint32_t v = 0;
        while (bits > have) {
            bits -= have;
            v += held << bits;
            held = src[at++];
            have = 8;
        }
        have -= bits;
        v += held >> have;
        held &= (1u << have) - 1u;
        return v;
    }
};

// ---- clamped table (entropy_encoding.hpp:51-122, 154-199) --------------------------------------
// Field widths climb 0 -> 1 -> 4 -> 8 -> 12 ... ; rung() numbers those widths 0,1,2,3,...
HOH_HD uint32_t next_width(uint32_t w) { return w == 0 ? 1u : (w == 1 ? 4u : w + 4u); }
HOH_HD uint32_t rung(uint32_t w) { return w == 0 ? 0u : (w == 1 ? 1u : w / 4u + 1u); }

struct ClampSet {
    uint16_t lo[16];
    uint16_t hi[16];
    uint32_t count;  // (prob_bits - 1) / 4 + 2
};

// One directional scan.  Walks symbols from one end, raising the running field width whenever a
// frequency does not fit, and records where each width first became necessary.  `bits_total`
// accumulates the running width per visited symbol (the size estimate).  Returns the stop index.
HOH_HD uint32_t clamp_walk(const uint32_t* freqs, uint32_t range, bool from_top, uint32_t prob_bits,
                           uint32_t count, uint16_t* marks, uint64_t* bits_total, uint32_t* last_width,
                           uint16_t unused_mark) {
    uint32_t width = 0, marked = 0;
    uint32_t i = from_top ? range - 1 : 0;
    for (;;) {
        if (!from_top && i >= range) break;
        while (freqs[i] >= (1u << width)) {
            uint32_t r = rung(width);
            if (r < count) marks[r] = (uint16_t)i;
            width = next_width(width);
            marked = r + 1;
        }
        if (width >= prob_bits) {
            width = prob_bits;
            *bits_total += width;
            break;
        }
        *bits_total += width;
        if (from_top) {
            if (i == 0) break;
            i--;
        } else {
            i++;
        }
    }
    for (; marked < count; marked++) marks[marked] = unused_mark;
    *last_width = width;
    return i;
}

HOH_HD uint32_t clamp_width_of(const ClampSet& c, uint32_t prob_bits, uint32_t sym) {  // :173-187
    uint32_t w = 0;
    if (c.lo[0] <= sym && c.hi[0] >= sym) w = 1;
    if (c.lo[1] <= sym && c.hi[1] >= sym) w = 4;
    for (uint32_t j = 2; j < c.count; j++)
        if (c.lo[j] <= sym && c.hi[j] >= sym) w = 4 * j;
    return w > prob_bits ? prob_bits : w;
}

// First part of what encode_entropy writes before the payload-length varint:
//   varint(range-1) varint(n) metadata
// plus the decisions the table needs: the clamp set and the table mode (1 = every frequency on maxbits
// bits, 2 = clamp pairs + variable-width frequencies).  Returns the bytes written; *stored_size gets
// entropy_encoding.hpp:45's size of the stored-mode alternative.
// representable != 0 (HOH_FIX_LONE): table mode 1 is not chosen when a frequency does not fit its maxbits-wide
// field (the reference writes it anyway and the table is lost, D6).
HOH_HD uint32_t plan_head(const uint32_t* freqs, uint32_t range, uint32_t n, uint32_t prob_bits, uint8_t* head,
                          uint32_t* stored_size, ClampSet* cs, uint32_t* table_mode, uint32_t representable = 0) {
    uint32_t at = 0;
    at = put_varint(head, at, range - 1);
    at = put_varint(head, at, n);
    uint32_t maxbits = bit_length(range - 1);
    *stored_size = at + 1 + (uint32_t)(((uint64_t)maxbits * n + 7) / 8);

    uint64_t raw_table_bytes = ((uint64_t)prob_bits * range + 7) / 8;  // :47
    cs->count = (prob_bits - 1) / 4 + 2;                               // :51
    for (int k = 0; k < 16; k++) cs->lo[k] = cs->hi[k] = 0;
    // :48-49 — 2*(maxbits-1) is int, the clamp count uint32_t: 32-bit unsigned product (range 1 wraps)
    uint64_t clamped_bits = (uint64_t)((uint32_t)(2 * ((int)maxbits - 1)) * cs->count) + 2ull * prob_bits;
    uint32_t w_up, w_down;
    uint32_t stop_up = clamp_walk(freqs, range, false, prob_bits, cs->count, cs->lo, &clamped_bits, &w_up,
                                  (uint16_t)(range - 1));
    uint32_t stop_down = clamp_walk(freqs, range, true, prob_bits, cs->count, cs->hi, &clamped_bits, &w_down, 0);
    // :121 — size_t arithmetic, wraps when the scans crossed
    clamped_bits += (uint64_t)w_down * ((uint64_t)stop_down - (uint64_t)stop_up - 1ull);
    uint64_t clamped_bytes = (clamped_bits + 7) / 8;
    *table_mode = raw_table_bytes < clamped_bytes ? 1u : 2u;  // :135 / :148
    if (representable && *table_mode == 1u)
        for (uint32_t s = 0; s < range; s++)
            if (freqs[s] >> maxbits) *table_mode = 2u;
    head[at++] = (uint8_t)((1u << 7) + (prob_bits << 2) + *table_mode);
    return at;
}

// Everything before the payload-length varint (plan_head + the table), bit-serially with the
// reference's packer: exact for every input including the over-wide fields of D6.  The kernels use it
// as is for those rare tables and a warp-parallel packer (same bytes when no field is over-wide) for
// the rest.
HOH_HD uint32_t build_head(const uint32_t* freqs, uint32_t range, uint32_t n, uint32_t prob_bits,
                           uint8_t* head, uint32_t* stored_size) {
    ClampSet cs;
    uint32_t mode;
    uint32_t at = plan_head(freqs, range, n, prob_bits, head, stored_size, &cs, &mode);
    uint32_t maxbits = bit_length(range - 1);
    BitSink sink;
    sink.open(head, at);
    if (mode == 1) {  // table mode 1, :135-147 (each freq on maxbits bits: D6)
        for (uint32_t s = 0; s < range; s++) sink.put(freqs[s], maxbits);
    } else {  // table mode 2, :148-200
        for (uint32_t j = 0; j < cs.count; j++) {
            sink.put(cs.lo[j], maxbits);
            sink.put(cs.hi[j], maxbits);
        }
        for (uint32_t s = 0; s < range; s++) sink.put(freqs[s], clamp_width_of(cs, prob_bits, s));
    }
    return sink.close();
}

// Parsed header of one stream (entropy_decoding.hpp:143-154).
struct StreamHead {
    uint32_t range, n, maxbits;
    uint32_t rans;        // metadata bit 7
    uint32_t prob_bits;   // 4- or 5-bit field depending on HOH_FIX_PROB_BITS5
    uint32_t table_mode;  // metadata & 3
    uint64_t body;        // byte offset just after the metadata byte
    uint32_t empty;       // n == 0 and the FIX_EMPTY flag: no metadata byte was consumed
};

template <typename Bytes>
HOH_HD StreamHead parse_head(const Bytes& src, uint64_t at, unsigned flags) {
    StreamHead h;
    h.range = get_varint(src, &at) + 1;
    h.n = get_varint(src, &at);
    h.maxbits = bit_length(h.range - 1);
    h.empty = 0;
    h.rans = h.prob_bits = h.table_mode = 0;
    if (h.n == 0 && (flags & 4u)) {
        h.empty = 1;
        h.body = at;
        return h;
    }
    uint32_t meta = src[at++];
    h.rans = meta >> 7;
    h.prob_bits = (flags & 1u) ? (meta & 0x7cu) >> 2 : (meta & 0x3cu) >> 2;
    h.table_mode = meta & 3u;
    h.body = at;
    return h;
}

// Table modes 1 and 2 (entropy_decoding.hpp:180-244): fills freqs[range]; returns the byte offset
// after the table.
template <typename Bytes>
HOH_HD uint64_t parse_table(const Bytes& src, const StreamHead& h, uint32_t* freqs) {
    BitSource<Bytes> bits{src, h.body, 0, 0};
    if (h.table_mode == 1) {
        for (uint32_t s = 0; s < h.range; s++) freqs[s] = bits.get(h.maxbits);
    } else {
        ClampSet cs;
        cs.count = (uint32_t)(((int)h.prob_bits - 1) / 4 + 2);  // :197, int arithmetic
        if (cs.count > 16) cs.count = 16;
        for (uint32_t j = 0; j < cs.count; j++) {
            cs.lo[j] = (uint16_t)bits.get(h.maxbits);
            cs.hi[j] = (uint16_t)bits.get(h.maxbits);
        }
        for (uint32_t s = 0; s < h.range; s++) freqs[s] = bits.get(clamp_width_of(cs, h.prob_bits, s));
    }
    return bits.at;
}

}  // namespace hohfmt
