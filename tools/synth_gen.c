/* Synthetic input generator for bench.py: integer-only, reproducible,
 * compressible RGB images and geometric symbol streams.  Input tooling, not part of the codec and
 * not part of any timed region.  Build: gcc -O2 -fPIC -shared -o libsynth.so synth_gen.c */
#include <stddef.h>
#include <stdint.h>

static inline uint64_t step(uint64_t* s) {
    uint64_t v = *s;
    v ^= v << 13;
    v ^= v >> 7;
    v ^= v << 17;
    return *s = v;
}

static inline int tri(int t, int period) {
    int m = t % (2 * period);
    return (m < period ? m : 2 * period - m) * 255 / period;
}

/* image `seed`: for y, x, c in raster order
 *   v = (tri(x+40c,97) + tri(y+24c,61) + tri(x+y,203)) / 3 + (next >> 61) - 4, clamped to 0..255 */
void synth_rgb(uint8_t* rgb, int w, int h, uint64_t seed) {
    uint64_t s = seed;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) {
                int noise = (int)(step(&s) >> 61) - 4;
                int v = (tri(x + 40 * c, 97) + tri(y + 24 * c, 61) + tri(x + y, 203)) / 3 + noise;
                rgb[((size_t)y * w + x) * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
}

/* images first_seed .. first_seed+count-1 back to back */
void synth_rgb_batch(uint8_t* rgb, int w, int h, uint64_t first_seed, size_t count) {
    for (size_t i = 0; i < count; i++) synth_rgb(rgb + i * (size_t)w * h * 3, w, h, first_seed + i);
}

/* geometric(p = 0.08) symbols clipped to 255, integer inverse CDF */
void synth_symbols(uint8_t* sym, size_t n, uint64_t seed) {
    uint32_t T[256];
    uint64_t t = 1u << 24;
    for (int k = 0; k < 256; k++) {
        T[k] = (uint32_t)t;
        t = t * 92 / 100;
    }
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) {
        uint32_t u = (uint32_t)(step(&s) >> 40);
        int k = 0;
        while (k < 255 && u < T[k + 1]) k++;
        sym[i] = (uint8_t)k;
    }
}
