// TEST INFRASTRUCTURE — not product code.
//
// Flat C interface over the UNMODIFIED reference (hohMiyazawa/hoh-ANS) so that
// tests can pin oracle/hoh_oracle.c against the real thing and bench.py can
// time the reference's own CPU implementation.  Nothing from the reference is
// copied: its headers are #included from where they lie (-I$(REF), normally
// /root/reference) at build time, and the resulting library goes to
// oracle/_ref/libhohref.so (git-ignored, travels to the GPU box with gpurun).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load the library built from this file.
//
// The reference prints to stdout from inside decode_entropy
// (entropy_decoding.hpp:257) and decode_layer (layer_decode.hpp:197,276); the
// printf macro below silences that without touching its arithmetic.

#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <assert.h>
#include <cstdio>   // libstdc++'s <cstdio> #undefs printf: pull every std header in BEFORE the macro
#include <cstdlib>
#include <cmath>
#include <fstream>

static inline int ref_noprintf(const char*, ...) { return 0; }
#define printf(...) ref_noprintf(__VA_ARGS__)
#define main choh_main
#include "choh.cpp"          // encode_tile, count_colours, palette_encode + all encode headers
#undef main
#include "layer_decode.hpp"  // decode_layer, unpredict_all (entropy_decoding.hpp already in)
#undef printf

extern "C" {

// entropy_encoding.hpp:8
size_t ref_encode_entropy(const uint16_t* symbols, size_t n, size_t range,
                          uint8_t* out, uint32_t prob_bits) {
    return encode_entropy(const_cast<uint16_t*>(symbols), n, range, out, prob_bits, 0);
}

// entropy_decoding.hpp:134 — copies the new[] result into `out` (capacity out_cap symbols).
// Returns the number of symbols; *byte_pointer is advanced exactly as the reference does (D8).
size_t ref_decode_entropy(const uint8_t* in, size_t in_size, size_t* byte_pointer,
                          uint16_t* out, size_t out_cap) {
    size_t n = 0;
    uint16_t* dec = decode_entropy(const_cast<uint8_t*>(in), in_size, byte_pointer, &n, 0);
    for (size_t i = 0; i < n && i < out_cap; i++) out[i] = dec[i];
    delete[] dec;
    return n;
}

// stattools.hpp:13
void ref_normalize_freqs(uint32_t* freqs, uint32_t* cum_freqs, size_t size, uint32_t target_total) {
    normalize_freqs(freqs, cum_freqs, size, target_total);
}

// channel.hpp:73
void ref_subtract_green(const uint8_t* rgb, size_t size, uint16_t* g, uint16_t* rg, uint16_t* bg) {
    subtract_green(const_cast<uint8_t*>(rgb), size, g, rg, bg);
}

// channel.hpp:63
void ref_channel_picker(const uint8_t* rgb, size_t size, int total, int target, uint16_t* out) {
    uint16_t* p = channel_picker(const_cast<uint8_t*>(rgb), size, total, target);
    memcpy(out, p, (size / total) * sizeof(uint16_t));
    delete[] p;
}

// prediction.hpp:6
size_t ref_predict_fastpath(const uint16_t* data, size_t size, int w, int h, int depth, uint16_t* out) {
    size_t n = 0;
    uint16_t* p = channelpredict_fastpath(const_cast<uint16_t*>(data), size, w, h, depth, &n);
    memcpy(out, p, n * sizeof(uint16_t));
    delete[] p;
    return n;
}

// prediction.hpp:46
size_t ref_predict_section(const uint16_t* data, size_t size, int w, int h, int depth,
                           size_t x_tiles, size_t y_tiles, int x, int y, uint16_t mask, uint16_t* out) {
    size_t n = 0;
    uint16_t* p = channelpredict_section(const_cast<uint16_t*>(data), size, w, h, depth,
                                         x_tiles, y_tiles, x, y, mask, &n);
    memcpy(out, p, n * sizeof(uint16_t));
    delete[] p;
    return n;
}

// prediction.hpp:153
void ref_predict_all(const uint16_t* data, size_t size, int w, int h, int depth,
                     int x_tiles, int y_tiles, const uint16_t* tile_map, uint16_t* out) {
    uint16_t* p = channelpredict_all(const_cast<uint16_t*>(data), size, w, h, depth,
                                     x_tiles, y_tiles, const_cast<uint16_t*>(tile_map));
    memcpy(out, p, size * sizeof(uint16_t));
    delete[] p;
}

// unprediction.hpp:6
void ref_unpredict_all(const uint16_t* resid, size_t size, int w, int h, int depth,
                       int x_tiles, int y_tiles, const uint16_t* tile_map,
                       const uint16_t* backref, uint16_t* out) {
    uint16_t* p = unpredict_all(const_cast<uint16_t*>(resid), size, w, h, depth, x_tiles, y_tiles,
                                const_cast<uint16_t*>(tile_map), const_cast<uint16_t*>(backref));
    memcpy(out, p, (size_t)w * h * sizeof(uint16_t));
    delete[] p;
}

// layer_encode.hpp:11
size_t ref_layer_encode(const uint16_t* plane, size_t size, int w, int h, int depth, size_t mode,
                        const uint8_t* nuke, uint8_t* out) {
    return layer_encode(const_cast<uint16_t*>(plane), size, w, h, depth, mode,
                        const_cast<uint8_t*>(nuke), out);
}

// layer_decode.hpp:128 — only usable on streams the reference can actually parse.
void ref_decode_layer(const uint8_t* in, size_t in_size, size_t byte_pointer, size_t w, size_t h,
                      uint8_t depth, const uint16_t* backref, uint8_t* out) {
    uint8_t* p = decode_layer(const_cast<uint8_t*>(in), in_size, byte_pointer, w, h, depth,
                              const_cast<uint16_t*>(backref));
    memcpy(out, p, w * h);
    delete[] p;
}

// choh.cpp:104
size_t ref_encode_tile(const uint8_t* rgb, size_t size, uint8_t* out, int w, int h, size_t mode) {
    return encode_tile(const_cast<uint8_t*>(rgb), size, out, w, h, mode);
}

// lz.hpp:6
size_t ref_find_lz_rgb(const uint8_t* rgb, size_t size, int w, int h, uint8_t* lz_symbols,
                       uint8_t* nukemap, int distance, int break_even_bonus) {
    return find_lz_rgb(const_cast<uint8_t*>(rgb), size, w, h, lz_symbols, nukemap, distance,
                       break_even_bonus);
}

// choh.cpp:17
int ref_count_colours(const uint8_t* rgb, size_t size) {
    return count_colours(const_cast<uint8_t*>(rgb), size);
}

// choh.cpp:394 — whole tool, file to file.  Returns main()'s exit code.
int ref_choh_main(const char* in_path, const char* out_path, int w, int h, int mode) {
    char ws[32], hs[32], ms[8], prog[8] = "choh";
    snprintf(ws, sizeof ws, "%d", w);
    snprintf(hs, sizeof hs, "%d", h);
    snprintf(ms, sizeof ms, "-s%d", mode);
    char* argv[7] = {prog, const_cast<char*>(in_path), const_cast<char*>(out_path), ws, hs, ms, 0};
    return choh_main(6, argv);
}

// rans64.hpp:262 loop with a caller-supplied static table (BASELINE config 4).
// Writes the payload words exactly as entropy_encoding.hpp:218-238 lays them out
// (ascending from the last-written word) and returns the byte count.
size_t ref_rans_encode_static(const uint16_t* symbols, size_t n, const uint32_t* freqs,
                              const uint32_t* cum_freqs, size_t range, uint32_t prob_bits,
                              uint8_t* out) {
    Rans64EncSymbol* esyms = new Rans64EncSymbol[range];
    for (size_t i = 0; i < range; i++) Rans64EncSymbolInit(&esyms[i], cum_freqs[i], freqs[i], prob_bits);
    size_t cap = n + 16;  // one word per symbol is an upper bound at prob_bits <= 31
    uint32_t* buf = new uint32_t[cap];
    uint32_t* ptr = buf + cap;
    Rans64State r;
    Rans64EncInit(&r);
    for (size_t i = n; i > 0; i--) Rans64EncPutSymbol(&r, &ptr, &esyms[symbols[i - 1]], prob_bits);
    Rans64EncFlush(&r, &ptr);
    size_t bytes = (size_t)(buf + cap - ptr) * 4;
    memcpy(out, ptr, bytes);
    delete[] buf;
    delete[] esyms;
    return bytes;
}

// rans64.hpp:107-142 loop with a caller-supplied static table (the body of
// entropy_decoding.hpp:262-276 without the header parse).
void ref_rans_decode_static(const uint8_t* in, size_t in_bytes, size_t n, const uint32_t* freqs,
                            const uint32_t* cum_freqs, size_t range, uint32_t prob_bits,
                            uint16_t* out) {
    uint16_t* cum2sym = new uint16_t[(size_t)1 << prob_bits];
    Rans64DecSymbol* dsyms = new Rans64DecSymbol[range];
    for (size_t s = 0; s < range; s++) {
        Rans64DecSymbolInit(&dsyms[s], cum_freqs[s], freqs[s]);
        for (uint32_t i = cum_freqs[s]; i < cum_freqs[s + 1]; i++) cum2sym[i] = (uint16_t)s;
    }
    uint32_t* words = new uint32_t[in_bytes / 4 + 4];  // aligned staging (+1 word of slack: the
    memset(words, 0, (in_bytes / 4 + 4) * 4);          //  decoder may read one word past the end)
    memcpy(words, in, in_bytes);
    uint32_t* ptr = words;
    Rans64State r;
    Rans64DecInit(&r, &ptr);
    for (size_t i = 0; i < n; i++) {
        uint32_t s = cum2sym[Rans64DecGet(&r, prob_bits)];
        out[i] = (uint16_t)s;
        Rans64DecAdvanceSymbol(&r, &ptr, &dsyms[s], prob_bits);
    }
    delete[] words;
    delete[] dsyms;
    delete[] cum2sym;
}

}  // extern "C"
