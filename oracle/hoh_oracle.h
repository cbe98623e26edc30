/* TEST INFRASTRUCTURE — CPU oracle for the hoh-ANS hot path.  NOT product code.
 *
 * A plain-C restatement of the reference algorithm (hohMiyazawa/hoh-ANS) for
 * the path the CUDA kernels replace: colour transform, prediction /
 * un-prediction, frequency statistics, stream header + table serialisation and
 * 64-bit rANS coding.  Every function cites the reference file:line it
 * follows.  Parity status: PINNED — tests/test_oracle_vs_ref.py checks every
 * function here against the real reference compiled from its own sources
 * (oracle/_ref/libhohref.so, see oracle/Makefile) and against the golden
 * vectors committed under tests/golden/ (which were generated from the real
 * reference by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this library.
 */
#ifndef HOH_ORACLE_H
#define HOH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes (the reference aborts via assert() where we return an error) */
#define ORC_OK 0
#define ORC_E_RANGE_GT_TOTAL 1 /* stattools.hpp:14 assert(target_total >= size) */
#define ORC_E_NO_DONOR 2       /* stattools.hpp:42 assert(best_steal != -1)     */
#define ORC_E_BAD_TABLE 3      /* decode: table storage mode 3 / inconsistent   */

/* decode flags: 0 = byte-for-byte reference behaviour (incl. its defects) */
#define ORC_FIX_PROB_BITS5 1u   /* D9: read prob_bits with the 5-bit mask 0x7C          */
#define ORC_FIX_ADVANCE 2u      /* D8: advance *byte_pointer past the rANS payload      */
#define ORC_FIX_EMPTY 4u        /* D2: a zero-symbol stream has no metadata byte        */
#define ORC_FIX_ALL 7u

/* ---- varint.hpp ------------------------------------------------------- */
size_t orc_write_varint(uint8_t* out, size_t pos, size_t value);        /* varint.hpp:29 */
size_t orc_read_varint(const uint8_t* in, size_t* pos);                  /* varint.hpp:6  */

/* ---- stattools.hpp ---------------------------------------------------- */
void orc_calc_cum_freqs(const uint32_t* freqs, uint32_t* cum, size_t size);               /* :6  */
int orc_normalize_freqs(uint32_t* freqs, uint32_t* cum, size_t size, uint32_t target);    /* :13 */

/* ---- entropy_encoding.hpp / entropy_decoding.hpp ---------------------- */
/* entropy_encoding.hpp:8.  Returns bytes written; *status (may be NULL) gets ORC_*. */
size_t orc_encode_entropy(const uint16_t* symbols, size_t n, size_t range, uint8_t* out,
                          uint32_t prob_bits, int* status);
/* entropy_decoding.hpp:134.  Writes up to out_cap symbols to out, returns the symbol count. */
size_t orc_decode_entropy(const uint8_t* in, size_t in_size, size_t* byte_pointer, uint16_t* out,
                          size_t out_cap, unsigned flags, int* status);
/* Header-only probe used by tests: fills range, n, metadata fields without decoding. */
void orc_peek_stream(const uint8_t* in, size_t pos, size_t* range, size_t* n, int* entropy_mode,
                     int* prob_bits5, int* table_mode);

/* rans64.hpp:262 loop / :107-142 loop with a caller-supplied table (config 4). */
size_t orc_rans_encode_static(const uint16_t* symbols, size_t n, const uint32_t* freqs,
                              const uint32_t* cum, size_t range, uint32_t prob_bits, uint8_t* out);
void orc_rans_decode_static(const uint8_t* in, size_t in_bytes, size_t n, const uint32_t* freqs,
                            const uint32_t* cum, size_t range, uint32_t prob_bits, uint16_t* out);

/* ---- channel.hpp ------------------------------------------------------ */
void orc_subtract_green(const uint8_t* rgb, size_t size, uint16_t* g, uint16_t* rg, uint16_t* bg); /* :73 */
void orc_channel_picker(const uint8_t* src, size_t size, int total, int target, uint16_t* out);    /* :63 */
/* Algebraic inverse of channel.hpp:73-79 (the reference has none: SURVEY §8.0 D4). */
void orc_add_green(const uint16_t* g, const uint16_t* rg, const uint16_t* bg, size_t pixels, uint8_t* rgb);

/* ---- predictor_operations.hpp (u16 forms) ----------------------------- */
uint16_t orc_midpoint(uint16_t a, uint16_t b);                /* :8  */
uint16_t orc_median(uint16_t a, uint16_t b, uint16_t c);      /* :37 */
uint16_t orc_average3(uint16_t a, uint16_t b, uint16_t c);    /* :66 */
uint16_t orc_paeth(uint16_t a, uint16_t b, uint16_t c);       /* :89 */

/* ---- prediction.hpp / unprediction.hpp -------------------------------- */
size_t orc_predict_fastpath(const uint16_t* data, int w, int h, int depth, uint16_t* out);          /* :6   */
size_t orc_predict_section(const uint16_t* data, int w, int h, int depth, size_t x_tiles,
                           size_t y_tiles, int x, int y, uint16_t mask, uint16_t* out);             /* :46  */
void orc_predict_all(const uint16_t* data, int w, int h, int depth, int x_tiles, int y_tiles,
                     const uint16_t* tile_map, uint16_t* out);                                      /* :153 */
void orc_unpredict_all(const uint16_t* resid, int w, int h, int depth, int x_tiles, int y_tiles,
                       const uint16_t* tile_map, const uint16_t* backref, uint16_t* out);           /* unprediction.hpp:6 */
/* Exact inverse of orc_predict_fastpath (pure MED, no last-row rule: SURVEY §8.0 D10). */
void orc_unpredict_fastpath(const uint16_t* resid, int w, int h, int depth, const uint16_t* backref,
                            uint16_t* out);

/* ---- layer_encode.hpp ------------------------------------------------- */
/* layer_encode.hpp:11, byte-for-byte including the stale-buffer behaviour (D7).
 * If trace != NULL it receives: [0]=winning residual-stream prob_bits as emitted (0 if stale),
 * [1]=x_tiles, [2]=y_tiles of the predictor grid, [3]=number of masks used. */
size_t orc_layer_encode(const uint16_t* plane, size_t size, int w, int h, int depth, size_t mode,
                        const uint8_t* nuke, uint8_t* out, int* trace);
/* Predictor search only (layer_encode.hpp:126-272): fills tile_map[x_tiles*y_tiles] (mask per
 * grid cell) and index_list (index into the 14 stock masks); returns number of cells. */
size_t orc_predictor_search(const uint16_t* plane, size_t size, int w, int h, int depth, size_t mode,
                            uint16_t* tile_map, uint8_t* index_list, uint16_t* final_resid);
extern const uint16_t orc_stock_masks[14];                                   /* layer_encode.hpp:159-175 */

/* ---- lz.hpp: the LZ match finder ------------------- */
/* lz.hpp:6-145.  rgb: size bytes (size = 3 * pixels).  nuke: pixels bytes, OR-ed with 1 where a match covers
 * the pixel (caller zeroes it, choh.cpp:118-121).  lz_out receives 0x03 + the 3 (distance <= 8) or 4 entropy
 * coded side streams; the return value is its length.  side / side_n (optional, may be NULL): the raw side
 * streams in the order since_last, length-4, back%256, back/256, each with capacity size/9 + 1. */
size_t orc_find_lz_rgb(const uint8_t* rgb, size_t size, int width, int distance, int break_even_bonus,
                       uint8_t* lz_out, uint8_t* nuke, uint8_t* const side[4], size_t side_n[4]);
/* choh.cpp:17-50: number of distinct colours, -1 if more than 256 */
int orc_count_colours(const uint8_t* rgb, size_t size);
/* choh.cpp:123-154: seek distance and break-even bonus encode_tile derives from the cruncher mode / colours */
void orc_lz_params(const uint8_t* rgb, size_t size, size_t mode, int* distance, int* bonus);

/* ---- synthetic inputs ----------------------------------- */
void orc_synth_rgb(uint8_t* rgb, int w, int h, uint64_t seed);
void orc_synth_symbols(uint8_t* sym, size_t n, uint64_t seed);   /* geometric(p=0.08), clipped to 255 */

#ifdef __cplusplus
}
#endif
#endif
