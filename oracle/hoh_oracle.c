/* TEST INFRASTRUCTURE — CPU oracle for the hoh-ANS hot path.  NOT product code.
 * See hoh_oracle.h for the contract.  All file:line citations are into the
 * reference tree (hohMiyazawa/hoh-ANS, mounted at /root/reference while
 * developing).  This is a restatement, not a copy: data structures, control
 * flow and naming are this repo's own; results are pinned bit-for-bit
 * against the compiled reference by tests/test_oracle_vs_ref.py.
 */
#include "hoh_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* varint.hpp                                                               */
/* ======================================================================== */

/* varint.hpp:29-45 — big-endian 7-bit groups, at most three; values >= 2^21 emit NOTHING (D5). */
size_t orc_write_varint(uint8_t* out, size_t pos, size_t v) {
    if (v < (1u << 7)) {
        out[pos++] = (uint8_t)v;
    } else if (v < (1u << 14)) {
        out[pos++] = (uint8_t)(0x80 | (v >> 7));
        out[pos++] = (uint8_t)(v & 0x7f);
    } else if (v < (1u << 21)) {
        out[pos++] = (uint8_t)(0x80 | (v >> 14));
        out[pos++] = (uint8_t)(0x80 | ((v >> 7) & 0x7f));
        out[pos++] = (uint8_t)(v & 0x7f);
    }
    return pos;
}

/* varint.hpp:6-27 — note the LAST byte is added whole (its top bit is not masked). */
size_t orc_read_varint(const uint8_t* in, size_t* pos) {
    size_t b0 = in[(*pos)++];
    if (!(b0 & 0x80)) return b0;
    size_t b1 = in[(*pos)++];
    if (!(b1 & 0x80)) return ((b0 & 0x7f) << 7) + b1;
    size_t b2 = in[(*pos)++];
    return ((b0 & 0x7f) << 14) + ((b1 & 0x7f) << 7) + b2;
}

/* MSB-first bit packer with the reference's arithmetic (varint.hpp:47-77 "stuffer").
 * The reference ADDS (not ORs) into the pending byte and never masks `value`, so a value wider
 * than `bits` carries into earlier bits of the byte (D6).  acc/avail mirror remainder /
 * bits_remaining. */
typedef struct {
    uint8_t* out;
    size_t pos;
    uint8_t acc;
    uint8_t avail; /* free bits in acc, 1..8 */
} bitw_t;

static void bitw_init(bitw_t* w, uint8_t* out, size_t pos) {
    w->out = out;
    w->pos = pos;
    w->acc = 0;
    w->avail = 8;
}

static void bitw_put_le8(bitw_t* w, uint32_t value, unsigned bits) { /* bits <= 8 */
    if (bits < w->avail) {
        w->acc = (uint8_t)(w->acc + (uint8_t)(value << (w->avail - bits)));
        w->avail = (uint8_t)(w->avail - bits);
    } else if (bits == w->avail) {
        w->out[w->pos++] = (uint8_t)(w->acc + (uint8_t)value);
        w->acc = 0;
        w->avail = 8;
    } else {
        unsigned spill = bits - w->avail;
        w->out[w->pos++] = (uint8_t)(w->acc + (uint8_t)(value >> spill));
        w->avail = (uint8_t)(8 - spill);
        w->acc = (uint8_t)((value << w->avail) & 0xff);
    }
}

static void bitw_put(bitw_t* w, uint32_t value, unsigned bits) {
    /* varint.hpp:65-70: a field wider than 8 bits (which always exceeds the free bits) is split
     * into value>>8 on bits-8 bits followed by the low byte on 8 bits, recursively. */
    if (bits > 8) {
        unsigned bytes_below = (bits - 1) / 8; /* number of whole low bytes */
        bitw_put_le8(w, value >> (8 * bytes_below), bits - 8 * bytes_below);
        while (bytes_below--) bitw_put_le8(w, (value >> (8 * bytes_below)) & 0xff, 8);
    } else {
        bitw_put_le8(w, value, bits);
    }
}

static void bitw_flush(bitw_t* w) { /* entropy_encoding.hpp:144-146 */
    if (w->avail != 8) w->out[w->pos++] = w->acc;
}

/* MSB-first bit reader (varint.hpp:79-106 "unstuffer"); starts with zero buffered bits. */
typedef struct {
    const uint8_t* in;
    size_t pos;
    uint8_t acc;
    uint8_t have;
} bitr_t;

static uint32_t bitr_get(bitr_t* r, unsigned bits) {
    uint32_t v = 0;
    while (bits > r->have) {
        bits -= r->have;
        v += (uint32_t)r->acc << bits;
        r->acc = r->in[r->pos++];
        r->have = 8;
    }
    r->have = (uint8_t)(r->have - bits);
    v += (uint32_t)(r->acc >> r->have);
    r->acc = (uint8_t)(r->acc & ((1u << r->have) - 1));
    return v;
}

/* ======================================================================== */
/* stattools.hpp                                                            */
/* ======================================================================== */

void orc_calc_cum_freqs(const uint32_t* freqs, uint32_t* cum, size_t size) { /* stattools.hpp:6-11 */
    uint32_t run = 0;
    for (size_t i = 0; i < size; i++) {
        cum[i] = run;
        run += freqs[i];
    }
    cum[size] = run;
}

/* stattools.hpp:13-70.  Rescale the cumulative counts to `target`, then make every used symbol
 * non-empty.  The reference shifts the cumulative boundaries between thief and donor by one
 * (:45-55); in frequency space that is exactly "donor -= 1, thief = 1", with the donor being the
 * lowest-index symbol of minimal current frequency > 1 (:33-41, strict <). */
int orc_normalize_freqs(uint32_t* freqs, uint32_t* cum, size_t size, uint32_t target) {
    if (target < size) return ORC_E_RANGE_GT_TOTAL;
    orc_calc_cum_freqs(freqs, cum, size);
    uint32_t total = cum[size];
    for (size_t i = 1; i <= size; i++) cum[i] = (uint32_t)(((uint64_t)target * cum[i]) / total);

    uint32_t* scaled = (uint32_t*)malloc(size * sizeof(uint32_t));
    for (size_t i = 0; i < size; i++) scaled[i] = cum[i + 1] - cum[i];
    for (size_t i = 0; i < size; i++) {
        if (!freqs[i] || scaled[i]) continue;
        uint32_t best = 0xffffffffu;
        long donor = -1;
        for (size_t j = 0; j < size; j++)
            if (scaled[j] > 1 && scaled[j] < best) {
                best = scaled[j];
                donor = (long)j;
            }
        if (donor < 0) {
            free(scaled);
            return ORC_E_NO_DONOR;
        }
        scaled[donor]--;
        scaled[i] = 1;
    }
    memcpy(freqs, scaled, size * sizeof(uint32_t));
    free(scaled);
    orc_calc_cum_freqs(freqs, cum, size);
    return ORC_OK;
}

/* ======================================================================== */
/* rans64.hpp                                                               */
/* ======================================================================== */

#define RANS_L (1ull << 31) /* rans64.hpp:59 */

/* One encoder step in the divide form (rans64.hpp:77-94).  SURVEY §7 H2 verified that the
 * reciprocal form (:262-278) used by encode_entropy is identical for every reachable
 * (x, freq, start, bits); tests re-check this through the compiled reference. */
static inline uint64_t rans_put(uint64_t x, uint32_t** pptr, uint32_t start, uint32_t freq, uint32_t bits) {
    uint64_t x_max = ((RANS_L >> bits) << 32) * freq;
    if (x >= x_max) {
        *--(*pptr) = (uint32_t)x;
        x >>= 32;
    }
    return ((x / freq) << bits) + (x % freq) + start;
}

/* entropy_encoding.hpp:218-227: reverse-order encode + flush into a descending word buffer;
 * returns pointer to the first payload word (payload = [ret, end)). */
static uint32_t* rans_encode_words(const uint16_t* sym, size_t n, const uint32_t* freqs,
                                   const uint32_t* cum, uint32_t bits, uint32_t* end) {
    uint32_t* p = end;
    uint64_t x = RANS_L; /* rans64.hpp:65 */
    for (size_t i = n; i-- > 0;) x = rans_put(x, &p, cum[sym[i]], freqs[sym[i]], bits);
    p -= 2; /* rans64.hpp:96-103 */
    p[0] = (uint32_t)x;
    p[1] = (uint32_t)(x >> 32);
    return p;
}

size_t orc_rans_encode_static(const uint16_t* symbols, size_t n, const uint32_t* freqs,
                              const uint32_t* cum, size_t range, uint32_t prob_bits, uint8_t* out) {
    (void)range;
    size_t cap = n + 16;
    uint32_t* buf = (uint32_t*)malloc(cap * sizeof(uint32_t));
    uint32_t* first = rans_encode_words(symbols, n, freqs, cum, prob_bits, buf + cap);
    size_t bytes = (size_t)(buf + cap - first) * 4;
    memcpy(out, first, bytes);
    free(buf);
    return bytes;
}

static inline uint32_t load_le32(const uint8_t* p) { /* entropy_decoding.hpp:269 unaligned u32 load */
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* entropy_decoding.hpp:262-276 with rans64.hpp:107-142; symbol lookup by binary search over cum
 * instead of the 2^prob_bits cum2sym array (same result by construction). */
static void rans_decode_symbols(const uint8_t* payload, size_t n, const uint32_t* freqs,
                                const uint32_t* cum, size_t range, uint32_t bits, uint16_t* out,
                                size_t out_cap) {
    const uint8_t* p = payload;
    uint64_t x = load_le32(p) | ((uint64_t)load_le32(p + 4) << 32);
    p += 8;
    uint64_t mask = (1ull << bits) - 1;
    for (size_t i = 0; i < n; i++) {
        uint32_t slot = (uint32_t)(x & mask);
        size_t lo = 0, hi = range; /* largest s with cum[s] <= slot and freq[s] > 0 */
        while (hi - lo > 1) {
            size_t mid = (lo + hi) / 2;
            if (cum[mid] <= slot) lo = mid; else hi = mid;
        }
        if (i < out_cap) out[i] = (uint16_t)lo;
        x = (uint64_t)freqs[lo] * (x >> bits) + slot - cum[lo];
        if (x < RANS_L) {
            x = (x << 32) | load_le32(p);
            p += 4;
        }
    }
}

void orc_rans_decode_static(const uint8_t* in, size_t in_bytes, size_t n, const uint32_t* freqs,
                            const uint32_t* cum, size_t range, uint32_t prob_bits, uint16_t* out) {
    /* The last step may read one word past the payload (the reference does too); stage with slack. */
    uint8_t* tmp = (uint8_t*)calloc(in_bytes + 8, 1);
    memcpy(tmp, in, in_bytes);
    rans_decode_symbols(tmp, n, freqs, cum, range, prob_bits, out, n);
    free(tmp);
}

/* ======================================================================== */
/* entropy_encoding.hpp                                                     */
/* ======================================================================== */

static unsigned bit_length(size_t v) { /* entropy_encoding.hpp:24-27 */
    unsigned b = 0;
    for (; v; v >>= 1) b++;
    return b;
}

/* Width ladder of the clamped table: 0 -> 1 -> 4 -> 8 -> 12 ... (entropy_encoding.hpp:61-75). */
static unsigned ladder_next(unsigned bits) { return bits == 0 ? 1 : (bits == 1 ? 4 : bits + 4); }
static unsigned ladder_slot(unsigned bits) { return bits == 0 ? 0 : (bits == 1 ? 1 : bits / 4 + 1); }

/* One directional clamp scan (entropy_encoding.hpp:56-86 ascending, :87-120 descending).
 * Returns the index where the scan stopped; adds the running width to *est per visited symbol.
 * Writes past clamp_number (possible only when one symbol owns the whole range and prob_bits is a
 * multiple of 4) land in padding in the reference build and are dropped here. */
static size_t clamp_scan(const uint32_t* freqs, size_t range, int descending, unsigned prob_bits,
                         unsigned clamp_number, uint16_t* clamps, uint64_t* est, unsigned* width_out,
                         uint16_t fill) {
    unsigned width = 0, filled = 0;
    size_t i = descending ? range - 1 : 0;
    for (;;) {
        if (!descending && i >= range) break;
        while (freqs[i] >= ((size_t)1 << width)) {
            unsigned slot = ladder_slot(width);
            if (slot < clamp_number) clamps[slot] = (uint16_t)i;
            width = ladder_next(width);
            filled = (width == 1) ? 1 : (width == 4 ? 2 : filled + 1);
        }
        if (width >= prob_bits) {
            width = prob_bits;
            *est += width;
            break;
        }
        *est += width;
        if (descending) {
            if (i == 0) break;
            i--;
        } else {
            i++;
        }
    }
    for (; filled < clamp_number; filled++) clamps[filled] = fill;
    *width_out = width;
    return i;
}

static unsigned clamped_width(const uint16_t* lo, const uint16_t* hi, unsigned clamp_number,
                              unsigned prob_bits, size_t i) { /* entropy_encoding.hpp:173-187 */
    unsigned w = 0;
    if (lo[0] <= i && hi[0] >= i) w = 1;
    if (lo[1] <= i && hi[1] >= i) w = 4;
    for (unsigned j = 2; j < clamp_number; j++)
        if (lo[j] <= i && hi[j] >= i) w = 4 * j;
    return w > prob_bits ? prob_bits : w;
}

size_t orc_encode_entropy(const uint16_t* symbols, size_t n, size_t range, uint8_t* out,
                          uint32_t prob_bits, int* status) {
    if (status) *status = ORC_OK;
    size_t pos = 0;
    pos = orc_write_varint(out, pos, range - 1); /* :43 (and :20 for the empty stream) */
    pos = orc_write_varint(out, pos, n);         /* :44 / :21 */
    if (n == 0) return pos;                      /* :19-23, no metadata byte (D2) */

    unsigned maxbits = bit_length(range - 1);
    uint32_t* freqs = (uint32_t*)calloc(range, sizeof(uint32_t));
    uint32_t* cum = (uint32_t*)calloc(range + 1, sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) freqs[symbols[i]]++; /* :37-39 */
    int st = orc_normalize_freqs(freqs, cum, range, 1u << prob_bits); /* :41 */
    if (st != ORC_OK) {
        if (status) *status = st;
        free(freqs);
        free(cum);
        return 0;
    }

    size_t stored_size = pos + 1 + (maxbits * n + 7) / 8;                 /* :45 */
    uint64_t raw_table = ((uint64_t)prob_bits * range + 7) / 8;           /* :47 */
    unsigned clamp_number = (prob_bits - 1) / 4 + 2;                      /* :51 */
    /* :48-49 — `2*(maxbits-1)` is int, the clamp count is uint32_t: the product is taken in 32-bit
     * unsigned arithmetic (matters only for range 1, where maxbits-1 == -1), then widened. */
    uint64_t clamped = (uint64_t)((uint32_t)(2 * ((int)maxbits - 1)) * (uint32_t)clamp_number) + 2ull * prob_bits;
    uint16_t lo[16], hi[16];
    memset(lo, 0, sizeof lo);
    memset(hi, 0, sizeof hi);
    unsigned w_up, w_down;
    size_t climb = clamp_scan(freqs, range, 0, prob_bits, clamp_number, lo, &clamped, &w_up, (uint16_t)(range - 1));
    size_t climb2 = clamp_scan(freqs, range, 1, prob_bits, clamp_number, hi, &clamped, &w_down, 0);
    (void)w_up;
    clamped += (uint64_t)w_down * (uint64_t)(climb2 - climb - 1); /* :121, unsigned wrap-around */
    clamped = (clamped + 7) / 8;                                  /* :122 */

    bitw_t bw;
    if (raw_table < clamped) { /* :135-147  table mode 1: every freq on maxbits bits (D6) */
        out[pos++] = (uint8_t)((1u << 7) + (prob_bits << 2) + 1);
        bitw_init(&bw, out, pos);
        for (size_t i = 0; i < range; i++) bitw_put(&bw, freqs[i], maxbits);
    } else { /* :148-200  table mode 2: clamp pairs, then variable-width freqs */
        out[pos++] = (uint8_t)((1u << 7) + (prob_bits << 2) + 2);
        bitw_init(&bw, out, pos);
        for (unsigned j = 0; j < clamp_number; j++) {
            bitw_put(&bw, lo[j], maxbits);
            bitw_put(&bw, hi[j], maxbits);
        }
        for (size_t i = 0; i < range; i++)
            bitw_put(&bw, freqs[i], clamped_width(lo, hi, clamp_number, prob_bits, i));
    }
    bitw_flush(&bw);
    pos = bw.pos;

    size_t cap = n + 16;
    uint32_t* words = (uint32_t*)malloc(cap * sizeof(uint32_t));
    uint32_t* first = rans_encode_words(symbols, n, freqs, cum, prob_bits, words + cap); /* :218-227 */
    size_t payload = (size_t)(words + cap - first) * 4;
    pos = orc_write_varint(out, pos, payload); /* :232 */
    memcpy(out + pos, first, payload);         /* :234-238, little-endian u32 stores */
    pos += payload;
    free(words);
    free(freqs);
    free(cum);

    if (stored_size < pos) { /* :244-267  stored mode: metadata 0, symbols on maxbits bits */
        pos = 0;
        pos = orc_write_varint(out, pos, range - 1);
        pos = orc_write_varint(out, pos, n);
        out[pos++] = 0;
        bitw_init(&bw, out, pos);
        for (size_t i = 0; i < n; i++) bitw_put(&bw, symbols[i], maxbits);
        bitw_flush(&bw);
        pos = bw.pos;
    }
    return pos;
}

/* ======================================================================== */
/* entropy_decoding.hpp                                                     */
/* ======================================================================== */

void orc_peek_stream(const uint8_t* in, size_t pos, size_t* range, size_t* n, int* entropy_mode,
                     int* prob_bits5, int* table_mode) {
    *range = orc_read_varint(in, &pos) + 1;
    *n = orc_read_varint(in, &pos);
    uint8_t meta = in[pos];
    *entropy_mode = meta >> 7;
    *prob_bits5 = (meta & 0x7c) >> 2;
    *table_mode = meta & 3;
}

size_t orc_decode_entropy(const uint8_t* in, size_t in_size, size_t* byte_pointer, uint16_t* out,
                          size_t out_cap, unsigned flags, int* status) {
    (void)in_size; /* entropy_decoding.hpp:136 — ignored by the reference too */
    if (status) *status = ORC_OK;
    size_t pos = *byte_pointer;
    size_t range = orc_read_varint(in, &pos) + 1; /* :143 */
    size_t n = orc_read_varint(in, &pos);         /* :144 */
    if (n == 0 && (flags & ORC_FIX_EMPTY)) {
        *byte_pointer = pos;
        return 0;
    }
    unsigned maxbits = bit_length(range - 1);
    uint8_t meta = in[pos++];                                         /* :151 */
    unsigned entropy_mode = meta >> 7;
    unsigned prob_bits = (flags & ORC_FIX_PROB_BITS5) ? (meta & 0x7c) >> 2 : (meta & 0x3c) >> 2; /* :153 (D9) */
    unsigned table_mode = meta & 3;

    bitr_t br = {in, pos, 0, 0};
    if (!entropy_mode) { /* :278-290 stored symbols */
        for (size_t i = 0; i < n; i++) {
            uint32_t v = bitr_get(&br, maxbits);
            if (i < out_cap) out[i] = (uint16_t)v;
        }
        *byte_pointer = br.pos;
        return n;
    }

    uint32_t* freqs = (uint32_t*)calloc(range, sizeof(uint32_t));
    uint32_t* cum = (uint32_t*)calloc(range + 1, sizeof(uint32_t));
    if (table_mode == 0) { /* :174-179 flat table */
        for (size_t i = 0; i < range; i++) freqs[i] = 1;
        int st = orc_normalize_freqs(freqs, cum, range, 1u << prob_bits);
        if (st != ORC_OK && status) *status = st;
    } else if (table_mode == 1) { /* :180-195 */
        for (size_t i = 0; i < range; i++) freqs[i] = bitr_get(&br, maxbits);
        orc_calc_cum_freqs(freqs, cum, range);
    } else if (table_mode == 2) { /* :196-244 */
        unsigned clamp_number = (unsigned)(((int)prob_bits - 1) / 4 + 2); /* :197, int arithmetic */
        uint16_t lo[16], hi[16];
        for (unsigned j = 0; j < clamp_number && j < 16; j++) {
            lo[j] = (uint16_t)bitr_get(&br, maxbits);
            hi[j] = (uint16_t)bitr_get(&br, maxbits);
        }
        for (size_t i = 0; i < range; i++)
            freqs[i] = bitr_get(&br, clamped_width(lo, hi, clamp_number, prob_bits, i));
        orc_calc_cum_freqs(freqs, cum, range);
    } else {
        if (status) *status = ORC_E_BAD_TABLE;
    }
    pos = br.pos;
    size_t payload = orc_read_varint(in, &pos); /* :256 */
    if (!status || *status == ORC_OK)
        rans_decode_symbols(in + pos, n, freqs, cum, range, prob_bits, out, out_cap); /* :262-276 */
    *byte_pointer = (flags & ORC_FIX_ADVANCE) ? pos + payload : pos;             /* (D8) */
    free(freqs);
    free(cum);
    return n;
}

/* ======================================================================== */
/* channel.hpp                                                              */
/* ======================================================================== */

void orc_subtract_green(const uint8_t* rgb, size_t size, uint16_t* g, uint16_t* rg, uint16_t* bg) {
    for (size_t p = 0; p < size / 3; p++) { /* channel.hpp:73-79 */
        int r = rgb[3 * p], gg = rgb[3 * p + 1], b = rgb[3 * p + 2];
        g[p] = (uint16_t)gg;
        rg[p] = (uint16_t)(r - gg + 256);
        bg[p] = (uint16_t)(b - gg + 256);
    }
}

void orc_channel_picker(const uint8_t* src, size_t size, int total, int target, uint16_t* out) {
    for (size_t p = 0; p < size / (size_t)total; p++) out[p] = src[p * total + target]; /* channel.hpp:63-71 */
}

void orc_add_green(const uint16_t* g, const uint16_t* rg, const uint16_t* bg, size_t pixels, uint8_t* rgb) {
    for (size_t p = 0; p < pixels; p++) { /* inverse of channel.hpp:75-77 */
        rgb[3 * p] = (uint8_t)((rg[p] + g[p] - 256) & 255);
        rgb[3 * p + 1] = (uint8_t)g[p];
        rgb[3 * p + 2] = (uint8_t)((bg[p] + g[p] - 256) & 255);
    }
}

/* ======================================================================== */
/* predictor_operations.hpp                                                 */
/* ======================================================================== */

uint16_t orc_midpoint(uint16_t a, uint16_t b) { /* :8-10, int arithmetic, division truncates toward 0 */
    return (uint16_t)((int)a + ((int)b - (int)a) / 2);
}

uint16_t orc_median(uint16_t a, uint16_t b, uint16_t c) { /* :37-60 */
    uint16_t lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : (c > hi ? hi : c);
}

uint16_t orc_average3(uint16_t a, uint16_t b, uint16_t c) { /* :66-68 */
    return (uint16_t)(((int)a + (int)b + (int)c) / 3);
}

uint16_t orc_paeth(uint16_t A, uint16_t B, uint16_t C) { /* :89-106 — NOT PNG's tie-breaking */
    int p = (int)A + (int)B - (int)C;
    int da = abs((int)A - p), db = abs((int)B - p), dc = abs((int)C - p);
    if (da < db) return da < dc ? A : C;
    return db < dc ? B : C;
}

/* The 16 candidate predictions, prediction.hpp:190-207 / unprediction.hpp:45-62.
 * section_order != 0 gives channelpredict_section's paeth(L,T,TL) argument order (:126). */
static void candidates(uint16_t L, uint16_t T, uint16_t TL, uint16_t TR, int section_order, uint16_t p[16]) {
    p[0] = L;
    p[1] = T;
    p[2] = TL;
    p[3] = TR;
    p[4] = orc_median(T, L, (uint16_t)(T + L - TL)); /* gradient wraps to u16 before the median (H6) */
    p[5] = orc_midpoint(L, T);
    p[6] = orc_midpoint(L, TL);
    p[7] = orc_midpoint(TL, T);
    p[8] = orc_midpoint(T, TR);
    p[9] = section_order ? orc_paeth(L, T, TL) : orc_paeth(L, TL, T);
    p[10] = orc_average3(L, L, TL);
    p[11] = orc_average3(L, TL, TL);
    p[12] = orc_average3(TL, TL, T);
    p[13] = orc_average3(TL, T, T);
    p[14] = orc_average3(T, T, TR);
    p[15] = orc_average3(T, TR, TR);
}

static int pick_best(uint16_t v, const uint16_t p[16], uint16_t mask, int centre) {
    /* prediction.hpp:138-146: lowest index wins ties; stays 0 if the mask is empty */
    int best = 0, best_err = centre * 2;
    for (int j = 0; j < 16; j++) {
        int err = abs((int)v - (int)p[j]);
        if (err < best_err && (mask & (1u << j))) {
            best_err = err;
            best = j;
        }
    }
    return best;
}

/* ======================================================================== */
/* prediction.hpp                                                           */
/* ======================================================================== */

size_t orc_predict_fastpath(const uint16_t* data, int w, int h, int depth, uint16_t* out) {
    int c = 1 << depth; /* prediction.hpp:6-44 */
    size_t k = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint16_t L = x ? data[y * w + x - 1] : (uint16_t)(c / 2);
            uint16_t T = y ? data[(y - 1) * w + x] : (uint16_t)(c / 2);
            uint16_t TL = (x && y) ? data[(y - 1) * w + x - 1] : (uint16_t)(c / 2);
            /* row 0: T comes from the c/2-initialised top_row and TL is the previous T, i.e. c/2;
             * column 0: L = TL = c/2 (forige/forige_TL reset each row). */
            if (!x) TL = (uint16_t)(c / 2);
            int pred = orc_median(T, L, (uint16_t)(T + L - TL));
            out[k++] = (uint16_t)(((int)data[y * w + x] - pred + c / 2 + c) % c);
        }
    return k;
}

size_t orc_predict_section(const uint16_t* data, int w, int h, int depth, size_t x_tiles,
                           size_t y_tiles, int x, int y, uint16_t mask, uint16_t* out) {
    if (mask == 0x0010 && x_tiles == 1 && y_tiles == 1 && x == 0 && y == 0) /* :59-68 */
        return orc_predict_fastpath(data, w, h, depth, out);
    int c = 1 << depth;
    int tw = (int)((w + x_tiles - 1) / x_tiles), th = (int)((h + y_tiles - 1) / y_tiles);
    int x0 = x * tw, y0 = y * th;
    int* bp = (int*)malloc(sizeof(int) * tw);
    uint16_t* top = (uint16_t*)malloc(sizeof(uint16_t) * tw);
    for (int i = 0; i < tw; i++) {
        bp[i] = 4; /* :76-79 */
        /* :85-94 — reads tw entries of the row above even when the cell is clipped at the right
         * image edge (the read then continues into the next image row). */
        top[i] = y ? data[y0 * w + x0 + i - w] : (uint16_t)(c / 2);
    }
    size_t k = 0;
    for (int ym = 0; ym < th && y0 + ym < h; ym++) {
        uint16_t left, left_top;
        if (x) { /* :97-105 */
            left = data[(y0 + ym) * w + x0 - 1];
            left_top = (ym || y) ? data[(y0 + ym - 1) * w + x0 - 1] : (uint16_t)(c / 2);
        } else {
            left = left_top = (uint16_t)(c / 2);
        }
        for (int xm = 0; xm < tw && x0 + xm < w; xm++) {
            uint16_t v = data[(y0 + ym) * w + x0 + xm];
            uint16_t p[16];
            candidates(left, top[xm], left_top, top[(xm + 1) % tw], 1, p); /* :115 TR wraps in the cell */
            int pred = orc_midpoint(p[bp[xm]], p[bp[(xm + tw - 1) % tw]]);  /* :134 */
            out[k++] = (uint16_t)(((int)v - pred + c / 2 + c) % c);
            left_top = top[xm];
            top[xm] = v;
            left = v;
            bp[xm] = pick_best(v, p, mask, c);
        }
    }
    free(bp);
    free(top);
    return k;
}

/* Shared raster walk of channelpredict_all (prediction.hpp:153-229) and unpredict_all
 * (unprediction.hpp:6-91): same neighbour rules, same best-predictor bookkeeping; only the
 * direction of the residual <-> value mapping differs. */
static void raster_walk(const uint16_t* in, int w, int h, int depth, int x_tiles, int y_tiles,
                        const uint16_t* tile_map, const uint16_t* backref, uint16_t* out, int inverse) {
    int c = 1 << depth;
    int tw = (w + x_tiles - 1) / x_tiles, th = (h + y_tiles - 1) / y_tiles;
    int* bp = (int*)malloc(sizeof(int) * w);
    uint16_t* top = (uint16_t*)malloc(sizeof(uint16_t) * w);
    for (int i = 0; i < w; i++) {
        bp[i] = 4;
        top[i] = (uint16_t)(c / 2);
    }
    size_t next_resid = 0;
    for (int y = 0; y < h; y++) {
        uint16_t left = (uint16_t)(c / 2), left_top = (uint16_t)(c / 2);
        for (int x = 0; x < w; x++) {
            int at = y * w + x;
            uint16_t p[16];
            candidates(left, top[x], left_top, top[(x + 1) % w], 0, p); /* TR of last column = top[0] (already this row) */
            int pred = orc_midpoint(p[bp[x]], p[bp[(x + w - 1) % w]]);
            uint16_t v;
            if (!inverse) {
                v = in[at];
                out[at] = (uint16_t)(((int)v - pred + c / 2 + c) % c); /* prediction.hpp:208 */
            } else if (backref && backref[at]) {
                v = out[at - backref[at]]; /* unprediction.hpp:63-65 */
                out[at] = v;
            } else {
                uint16_t t = (uint16_t)((int)in[next_resid++] - c - c / 2 + pred); /* unprediction.hpp:67 */
                v = (uint16_t)(t % c);
                out[at] = v;
            }
            left_top = top[x];
            top[x] = v;
            left = v;
            if (y + 1 < h) /* prediction.hpp:216-225: mask of the cell BELOW; last row forces 0 */
                bp[x] = pick_best(v, p, tile_map[((y + 1) / th) * x_tiles + x / tw], c);
            else
                bp[x] = 0;
        }
    }
    free(bp);
    free(top);
}

void orc_predict_all(const uint16_t* data, int w, int h, int depth, int x_tiles, int y_tiles,
                     const uint16_t* tile_map, uint16_t* out) {
    raster_walk(data, w, h, depth, x_tiles, y_tiles, tile_map, NULL, out, 0);
}

void orc_unpredict_all(const uint16_t* resid, int w, int h, int depth, int x_tiles, int y_tiles,
                       const uint16_t* tile_map, const uint16_t* backref, uint16_t* out) {
    raster_walk(resid, w, h, depth, x_tiles, y_tiles, tile_map, backref, out, 1);
}

void orc_unpredict_fastpath(const uint16_t* resid, int w, int h, int depth, const uint16_t* backref,
                            uint16_t* out) {
    int c = 1 << depth;
    size_t k = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int at = y * w + x;
            if (backref && backref[at]) {
                out[at] = out[at - backref[at]];
                continue;
            }
            uint16_t L = x ? out[at - 1] : (uint16_t)(c / 2);
            uint16_t T = y ? out[at - w] : (uint16_t)(c / 2);
            uint16_t TL = (x && y) ? out[at - w - 1] : (uint16_t)(c / 2);
            int pred = orc_median(T, L, (uint16_t)(T + L - TL));
            out[at] = (uint16_t)(((int)resid[k++] + pred - c / 2) & (c - 1));
        }
}

/* ======================================================================== */
/* layer_encode.hpp                                                         */
/* ======================================================================== */

const uint16_t orc_stock_masks[14] = {/* layer_encode.hpp:159-175 */
                                      0x0001, 0x0002, 0x0020, 0x0010, 0xffbf, 0x0003, 0xfffd,
                                      0xfffb, 0xfff7, 0xffef, 0xffdf, 0xff7f, 0xfdff, 0xffff};

#define GRID 40 /* layer_encode.hpp:124 */

/* layer_encode.hpp:133-144 (and :217-225): +1-smoothed histogram -> per-symbol cost in bits. */
static void cost_table(const uint16_t* resid, size_t size, size_t range, double* cost) {
    int* f = (int*)malloc(range * sizeof(int));
    for (size_t i = 0; i < range; i++) f[i] = 1;
    for (size_t i = 0; i < size; i++) f[resid[i]]++;
    for (size_t i = 0; i < range; i++) cost[i] = -log2((double)f[i] / (double)size);
    free(f);
}

/* layer_encode.hpp:176-203: per grid cell, first strictly-cheapest of the first min(14,5*mode)
 * stock masks; cost summed in raster order in double precision. */
static void search_pass(const uint16_t* plane, int w, int h, int depth, int xt, int yt, size_t mode,
                        const double* cost, uint16_t* tile_map, uint8_t* index_list) {
    uint16_t* cell = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(GRID + 1) * (GRID + 1));
    for (int t = 0; t < xt * yt; t++) {
        double best = 99999999999.0;
        for (size_t m = 0; m < 14 && m < mode * 5; m++) {
            size_t cnt = orc_predict_section(plane, w, h, depth, (size_t)xt, (size_t)yt, t % xt, t / xt,
                                             orc_stock_masks[m], cell);
            double sum = 0;
            for (size_t k = 0; k < cnt; k++) sum += cost[cell[k]];
            if (sum < best) {
                best = sum;
                tile_map[t] = orc_stock_masks[m];
                index_list[t] = (uint8_t)m;
            }
        }
    }
    free(cell);
}

size_t orc_predictor_search(const uint16_t* plane, size_t size, int w, int h, int depth, size_t mode,
                            uint16_t* tile_map, uint8_t* index_list, uint16_t* final_resid) {
    size_t range = (size_t)1 << depth;
    int xt = (w + GRID - 1) / GRID, yt = (h + GRID - 1) / GRID;
    double* cost = (double*)malloc(range * sizeof(double));
    uint16_t* resid = (uint16_t*)malloc(size * sizeof(uint16_t));
    orc_predict_fastpath(plane, w, h, depth, resid);
    cost_table(resid, size, range, cost);
    search_pass(plane, w, h, depth, xt, yt, mode, cost, tile_map, index_list);
    orc_predict_all(plane, w, h, depth, xt, yt, tile_map, resid);
    if (mode > 2) { /* :215-272 refinement */
        cost_table(resid, size, range, cost);
        search_pass(plane, w, h, depth, xt, yt, mode, cost, tile_map, index_list);
        orc_predict_all(plane, w, h, depth, xt, yt, tile_map, resid);
    }
    if (final_resid) memcpy(final_resid, resid, size * sizeof(uint16_t));
    free(cost);
    free(resid);
    return (size_t)xt * yt;
}

static size_t compact(const uint16_t* resid, const uint8_t* nuke, size_t size, uint16_t* dst) {
    size_t k = 0; /* layer_encode.hpp:93-99 / :328-333 */
    for (size_t i = 0; i < size; i++)
        if (nuke[i] == 0) dst[k++] = resid[i];
    return k;
}

size_t orc_layer_encode(const uint16_t* plane, size_t size, int w, int h, int depth, size_t mode,
                        const uint8_t* nuke, uint8_t* out, int* trace) {
    size_t o = 0;
    size_t best_size = ((size_t)depth * size + ((size_t)depth * size) % 8 + 1024) / 8; /* :22 */
    out[o++] = 0x10;                                                                  /* :57 */

    uint16_t* resid = (uint16_t*)malloc(size * sizeof(uint16_t));
    uint16_t* dense = (uint16_t*)malloc(size * sizeof(uint16_t));
    orc_predict_fastpath(plane, w, h, depth, resid); /* :63-75 */
    size_t dense_n = compact(resid, nuke, size, dense);

    size_t cap = 1024 + ((size_t)1 << depth) * 2 + (dense_n * (size_t)depth + 7) / 8 + (size_t)4 * dense_n + 64;
    /* two scratch buffers that trade places (:103-120); `kept` is what finally gets emitted */
    uint8_t* work = (uint8_t*)calloc(cap, 1);
    uint8_t* kept = (uint8_t*)calloc(cap, 1);
    int kept_bits = 0;
    size_t sz = orc_encode_entropy(dense, dense_n, (size_t)1 << depth, work, 15, NULL); /* :106 */
    if (sz < best_size) {
        best_size = sz;
        uint8_t* t = kept; kept = work; work = t;
        kept_bits = 15;
    }

    int xt = (w + GRID - 1) / GRID, yt = (h + GRID - 1) / GRID;
    int n_used = 0;
    if (mode && (xt > 1 || yt > 1)) { /* :126-319 */
        size_t cells = (size_t)xt * yt;
        uint16_t* tile_map = (uint16_t*)malloc(cells * sizeof(uint16_t));
        uint8_t* index_list = (uint8_t*)malloc(cells);
        orc_predictor_search(plane, size, w, h, depth, mode, tile_map, index_list, resid);
        out[o++] = (uint8_t)(xt - 1); /* :276-277 */
        out[o++] = (uint8_t)(yt - 1);
        uint8_t used[14] = {0}, remap[14] = {0};
        for (size_t t = 0; t < cells; t++) used[index_list[t]] = 1;
        for (int m = 0; m < 14; m++) n_used += used[m];
        out[o++] = (uint8_t)n_used; /* :291 */
        uint8_t next = 0;
        for (int m = 0; m < 14; m++)
            if (used[m]) { /* :292-304 */
                out[o++] = (uint8_t)(orc_stock_masks[m] >> 8);
                out[o++] = (uint8_t)(orc_stock_masks[m] & 0xff);
                remap[m] = next++;
            }
        uint16_t* idx16 = (uint16_t*)malloc(cells * sizeof(uint16_t));
        for (size_t t = 0; t < cells; t++) idx16[t] = remap[index_list[t]];
        o += orc_encode_entropy(idx16, cells, (size_t)n_used, out + o, 8, NULL); /* :308-317 */
        free(idx16);
        free(tile_map);
        free(index_list);
    } else { /* :320-325 */
        xt = yt = 1;
        out[o++] = 0;
        out[o++] = 0;
        out[o++] = 0x00;
        out[o++] = 0x10;
    }

    if (mode) { /* :326-392 prob_bits search; NB only the 17-19 / 14-12 trials ever swap buffers (D7) */
        dense_n = compact(resid, nuke, size, dense);
        size_t range = (size_t)1 << depth;
        size_t s16 = orc_encode_entropy(dense, dense_n, range, work, 16, NULL);
        size_t s15 = orc_encode_entropy(dense, dense_n, range, work, 15, NULL);
        int up = s16 < s15;
        size_t first = up ? s16 : s15;
        if (first < best_size) {
            best_size = first; /* size updated, bytes NOT kept */
            kept_bits = 0;
        }
        for (int k = 0; k < 3; k++) {
            uint32_t bits = up ? (uint32_t)(17 + k) : (uint32_t)(14 - k);
            size_t s = orc_encode_entropy(dense, dense_n, range, work, bits, NULL);
            if (s < best_size) {
                best_size = s;
                uint8_t* t = kept; kept = work; work = t;
                kept_bits = (int)bits;
            }
        }
    }
    memcpy(out + o, kept, best_size); /* :396-398 */
    o += best_size;
    if (trace) {
        trace[0] = kept_bits;
        trace[1] = xt;
        trace[2] = yt;
        trace[3] = n_used;
    }
    free(work);
    free(kept);
    free(resid);
    free(dense);
    return o;
}

/* ======================================================================== */
/* lz.hpp — LZ match finder over RGB pixels                                 */
/* ======================================================================== */

/* Number of consecutive pixels, at most 259, for which pixel (at + k) equals pixel (at - back + k); the run
 * also stops at the end of the image (lz.hpp:37-46 and :55-64 — the two loops differ only in how the end
 * bound is written, and for size % 3 == 0 the bounds coincide). */
static int lz_run(const uint8_t* rgb, size_t size, size_t at, size_t back_bytes) {
    int k = 0;
    while (k < 259) {
        size_t a = at + (size_t)k * 3;
        if (a + 2 >= size) break;
        const uint8_t* p = rgb + a;
        const uint8_t* q = p - back_bytes;
        if (p[0] != q[0] || p[1] != q[1] || p[2] != q[2]) break;
        k++;
    }
    return k;
}

size_t orc_find_lz_rgb(const uint8_t* rgb, size_t size, int width, int distance, int bonus, uint8_t* lz_out,
                       uint8_t* nuke, uint8_t* const side_out[4], size_t side_n_out[4]) {
    const size_t cap = size / 9 + 1;
    uint8_t* side[4];
    size_t cnt[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; k++) side[k] = (uint8_t*)malloc(cap);
    const long near_limit = 1L << distance; /* :20 */
    const int wide = distance > 8;
    int gap = 0; /* pixels since the last match, wraps at 255 (:77-83) */
    for (size_t at = 0; at < size; at += 3) {
        int best_len = 0;
        long best_back = -1;
        /* :34-52 every distance 1..2^distance; the first (smallest) distance with the longest run wins */
        for (long back = 1; back <= near_limit && (long)(int)at - back * 3 >= 0; back++) {
            int len = lz_run(rgb, size, at, (size_t)back * 3);
            if (len > best_len) {
                best_len = len;
                best_back = back;
                if (len == 259) break;
            }
        }
        /* :53-74 whole rows up: multiples of the width up to 65536 */
        if (best_len < 259 && wide)
            for (long back = width; back <= 65536 && (long)(int)at - back * 3 >= 0; back += width) {
                int len = lz_run(rgb, size, at, (size_t)back * 3);
                if (len > best_len) {
                    best_len = len;
                    best_back = back;
                    if (len == 259) break; /* nothing can be longer; the reference's odd restart (:70) finds nothing new */
                }
            }
        if (best_len < 4 + bonus) { /* :75-83 */
            if (++gap == 255) {
                gap = 0;
                side[0][cnt[0]++] = 255;
            }
            continue;
        }
        side[0][cnt[0]++] = (uint8_t)gap; /* :85-91 */
        if (wide) side[3][cnt[3]++] = (uint8_t)(best_back / 256);
        side[2][cnt[2]++] = (uint8_t)(best_back % 256);
        side[1][cnt[1]++] = (uint8_t)(best_len - 4);
        gap = 0;
        for (int k = 0; k < best_len; k++) nuke[at / 3 + (size_t)k] = 1; /* :92-94 */
        at += (size_t)(best_len - 1) * 3;                                 /* :95 */
    }
    /* :100-142: tag byte, then each side stream through encode_entropy (8-bit overload, range 256, 10 bits) */
    size_t o = 0;
    lz_out[o++] = 0x03;
    uint16_t* wide_sym = (uint16_t*)malloc(cap * sizeof(uint16_t));
    for (int k = 0; k < (wide ? 4 : 3); k++) {
        for (size_t i = 0; i < cnt[k]; i++) wide_sym[i] = side[k][i];
        o += orc_encode_entropy(wide_sym, cnt[k], 256, lz_out + o, 10, NULL);
    }
    free(wide_sym);
    for (int k = 0; k < 4; k++) {
        if (side_out && side_out[k]) memcpy(side_out[k], side[k], cnt[k]);
        if (side_n_out) side_n_out[k] = cnt[k];
        free(side[k]);
    }
    return o;
}

/* choh.cpp:17-50 */
int orc_count_colours(const uint8_t* rgb, size_t size) {
    uint8_t pal[257][3];
    int n = 0;
    for (size_t i = 0; i + 2 < size; i += 3) {
        int j = 0;
        while (j < n && (pal[j][0] != rgb[i] || pal[j][1] != rgb[i + 1] || pal[j][2] != rgb[i + 2])) j++;
        if (j < n) continue;
        memcpy(pal[n], rgb + i, 3);
        if (++n == 257) return -1;
    }
    return n;
}

/* choh.cpp:123-154 */
void orc_lz_params(const uint8_t* rgb, size_t size, size_t mode, int* distance, int* bonus) {
    static const int dist_of_mode[5] = {6, 10, 11, 12, 14};
    *distance = mode <= 4 ? dist_of_mode[mode] : 6;
    int colours = orc_count_colours(rgb, size);
    *bonus = 0;
    if (colours != -1) *bonus = colours <= 4 ? 32 : colours <= 8 ? 20 : colours <= 16 ? 10 : colours <= 32 ? 2 : 0;
}

/* ======================================================================== */
/* synthetic inputs                                          */
/* ======================================================================== */

static inline uint64_t xorshift64(uint64_t* s) {
    uint64_t v = *s;
    v ^= v << 13;
    v ^= v >> 7;
    v ^= v << 17;
    return *s = v;
}

static inline int tri(int t, int P) {
    int m = t % (2 * P);
    return (m < P ? m : 2 * P - m) * 255 / P;
}

void orc_synth_rgb(uint8_t* rgb, int w, int h, uint64_t seed) {
    uint64_t s = seed;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) {
                uint64_t r = xorshift64(&s);
                int v = (tri(x + 40 * c, 97) + tri(y + 24 * c, 61) + tri(x + y, 203)) / 3 + (int)(r >> 61) - 4;
                rgb[((size_t)y * w + x) * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
}

void orc_synth_symbols(uint8_t* sym, size_t n, uint64_t seed) {
    /* integer inverse-CDF of geometric(p = 0.08): T[k] = floor(2^24 * 0.92^k) by repeated *92/100 */
    uint32_t T[256];
    uint64_t t = 1u << 24;
    for (int k = 0; k < 256; k++) {
        T[k] = (uint32_t)t;
        t = t * 92 / 100;
    }
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) {
        uint32_t u = (uint32_t)(xorshift64(&s) >> 40);
        int k = 0;
        while (k < 255 && u < T[k + 1]) k++;
        sym[i] = (uint8_t)k;
    }
}
