"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when it was built, the real
reference behind its flat C wrapper (oracle/_ref/libhohref.so).

TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libhohref.so")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
sz = C.c_size_t


def build_oracle(force=False):
    """Compile oracle/ (and oracle/_ref when /root/reference is present)."""
    src = [os.path.join(ORACLE_DIR, f) for f in ("hoh_oracle.c", "hoh_oracle.h")]
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_src = os.path.join(ORACLE_DIR, "ref_wrap.cpp")
    if os.path.exists("/root/reference/choh.cpp"):
        if force or not os.path.exists(REF_SO) or os.path.getmtime(ref_src) > os.path.getmtime(REF_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_write_varint.restype = sz
        L.orc_write_varint.argtypes = [u8p, sz, sz]
        L.orc_normalize_freqs.restype = C.c_int
        L.orc_normalize_freqs.argtypes = [u32p, u32p, sz, C.c_uint32]
        L.orc_encode_entropy.restype = sz
        L.orc_encode_entropy.argtypes = [u16p, sz, sz, u8p, C.c_uint32, C.POINTER(C.c_int)]
        L.orc_decode_entropy.restype = sz
        L.orc_decode_entropy.argtypes = [u8p, sz, C.POINTER(sz), u16p, sz, C.c_uint, C.POINTER(C.c_int)]
        L.orc_rans_encode_static.restype = sz
        L.orc_rans_encode_static.argtypes = [u16p, sz, u32p, u32p, sz, C.c_uint32, u8p]
        L.orc_rans_decode_static.restype = None
        L.orc_rans_decode_static.argtypes = [u8p, sz, sz, u32p, u32p, sz, C.c_uint32, u16p]
        L.orc_subtract_green.restype = None
        L.orc_subtract_green.argtypes = [u8p, sz, u16p, u16p, u16p]
        L.orc_channel_picker.restype = None
        L.orc_channel_picker.argtypes = [u8p, sz, C.c_int, C.c_int, u16p]
        L.orc_add_green.restype = None
        L.orc_add_green.argtypes = [u16p, u16p, u16p, sz, u8p]
        L.orc_predict_fastpath.restype = sz
        L.orc_predict_fastpath.argtypes = [u16p, C.c_int, C.c_int, C.c_int, u16p]
        L.orc_predict_section.restype = sz
        L.orc_predict_section.argtypes = [u16p, C.c_int, C.c_int, C.c_int, sz, sz, C.c_int, C.c_int,
                                          C.c_uint16, u16p]
        L.orc_predict_all.restype = None
        L.orc_predict_all.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u16p, u16p]
        L.orc_unpredict_all.restype = None
        L.orc_unpredict_all.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u16p,
                                        C.c_void_p, u16p]
        L.orc_unpredict_fastpath.restype = None
        L.orc_unpredict_fastpath.argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_void_p, u16p]
        L.orc_layer_encode.restype = sz
        L.orc_layer_encode.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, sz, u8p, u8p,
                                       C.POINTER(C.c_int)]
        L.orc_predictor_search.restype = sz
        L.orc_predictor_search.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, sz, u16p, u8p, C.c_void_p]
        L.orc_find_lz_rgb.restype = sz
        L.orc_find_lz_rgb.argtypes = [u8p, sz, C.c_int, C.c_int, C.c_int, u8p, u8p, C.POINTER(C.c_void_p),
                                      C.POINTER(sz)]
        L.orc_count_colours.restype = C.c_int
        L.orc_count_colours.argtypes = [u8p, sz]
        L.orc_lz_params.restype = None
        L.orc_lz_params.argtypes = [u8p, sz, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_synth_rgb.restype = None
        L.orc_synth_rgb.argtypes = [u8p, C.c_int, C.c_int, C.c_uint64]
        L.orc_synth_symbols.restype = None
        L.orc_synth_symbols.argtypes = [u8p, sz, C.c_uint64]
        for name in ("orc_midpoint", "orc_median", "orc_average3", "orc_paeth"):
            f = getattr(L, name)
            f.restype = C.c_uint16
            f.argtypes = [C.c_uint16] * (2 if name == "orc_midpoint" else 3)
        _oracle = L
    return _oracle


def have_ref():
    build_oracle()
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        build_oracle()
        L = C.CDLL(REF_SO)
        L.ref_encode_entropy.restype = sz
        L.ref_encode_entropy.argtypes = [u16p, sz, sz, u8p, C.c_uint32]
        L.ref_decode_entropy.restype = sz
        L.ref_decode_entropy.argtypes = [u8p, sz, C.POINTER(sz), u16p, sz]
        L.ref_normalize_freqs.restype = None
        L.ref_normalize_freqs.argtypes = [u32p, u32p, sz, C.c_uint32]
        L.ref_subtract_green.restype = None
        L.ref_subtract_green.argtypes = [u8p, sz, u16p, u16p, u16p]
        L.ref_channel_picker.restype = None
        L.ref_channel_picker.argtypes = [u8p, sz, C.c_int, C.c_int, u16p]
        L.ref_predict_fastpath.restype = sz
        L.ref_predict_fastpath.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, u16p]
        L.ref_predict_section.restype = sz
        L.ref_predict_section.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, sz, sz, C.c_int, C.c_int,
                                          C.c_uint16, u16p]
        L.ref_predict_all.restype = None
        L.ref_predict_all.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u16p, u16p]
        L.ref_unpredict_all.restype = None
        L.ref_unpredict_all.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u16p,
                                        u16p, u16p]
        L.ref_layer_encode.restype = sz
        L.ref_layer_encode.argtypes = [u16p, sz, C.c_int, C.c_int, C.c_int, sz, u8p, u8p]
        L.ref_decode_layer.restype = None
        L.ref_decode_layer.argtypes = [u8p, sz, sz, sz, sz, C.c_uint8, u16p, u8p]
        L.ref_encode_tile.restype = sz
        L.ref_encode_tile.argtypes = [u8p, sz, u8p, C.c_int, C.c_int, sz]
        L.ref_find_lz_rgb.restype = sz
        L.ref_find_lz_rgb.argtypes = [u8p, sz, C.c_int, C.c_int, u8p, u8p, C.c_int, C.c_int]
        L.ref_count_colours.restype = C.c_int
        L.ref_count_colours.argtypes = [u8p, sz]
        L.ref_choh_main.restype = C.c_int
        L.ref_choh_main.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.ref_rans_encode_static.restype = sz
        L.ref_rans_encode_static.argtypes = [u16p, sz, u32p, u32p, sz, C.c_uint32, u8p]
        L.ref_rans_decode_static.restype = None
        L.ref_rans_decode_static.argtypes = [u8p, sz, sz, u32p, u32p, sz, C.c_uint32, u16p]
        _ref = L
    return _ref


# ---------------------------------------------------------------------------------------------
# numpy-level helpers (oracle side)
# ---------------------------------------------------------------------------------------------

def out_capacity(n, range_):
    return 2048 + 4 * range_ + 4 * n


def orc_encode_entropy(symbols, range_, prob_bits):
    symbols = np.ascontiguousarray(symbols, dtype=np.uint16)
    out = np.zeros(out_capacity(len(symbols), range_), np.uint8)
    st = C.c_int(0)
    n = oracle().orc_encode_entropy(symbols, len(symbols), range_, out, prob_bits, C.byref(st))
    return out[:n].copy(), st.value


def ref_encode_entropy(symbols, range_, prob_bits):
    symbols = np.ascontiguousarray(symbols, dtype=np.uint16)
    out = np.zeros(out_capacity(len(symbols), range_), np.uint8)
    n = ref().ref_encode_entropy(symbols, len(symbols), range_, out, prob_bits)
    return out[:n].copy()


def _padded_for_decode(stream, pos):
    """The reference's decoder (and the oracle, which mirrors it) trusts the stream: with a table the format could
    not represent the state collapses and it reads one 32-bit word per remaining symbol past the
    payload.  Give it zeros to read instead of whatever follows the array in memory: 4 bytes per declared
    symbol beyond the stream's own bytes."""
    n = 0
    try:
        at = pos
        for _ in range(2):  # varint(range - 1), varint(n): varint.hpp:6-27
            b0 = int(stream[at]); at += 1
            v = b0
            if b0 & 0x80:
                b1 = int(stream[at]); at += 1
                v = ((b0 & 0x7f) << 7) + b1
                if b1 & 0x80:
                    v = ((b0 & 0x7f) << 14) + ((b1 & 0x7f) << 7) + int(stream[at]); at += 1
            n = v
    except IndexError:
        n = 0
    return np.concatenate([stream, np.zeros(4 * n + 4096, np.uint8)])


def orc_decode_entropy(stream, pos=0, flags=7, cap=1 << 22):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    padded = _padded_for_decode(stream, pos)
    out = np.zeros(cap, np.uint16)
    bp = sz(pos)
    st = C.c_int(0)
    n = oracle().orc_decode_entropy(padded, len(stream), C.byref(bp), out, cap, flags, C.byref(st))
    return out[:n].copy(), bp.value, st.value


def ref_decode_entropy(stream, pos=0, cap=1 << 22):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    padded = _padded_for_decode(stream, pos)
    out = np.zeros(cap, np.uint16)
    bp = sz(pos)
    n = ref().ref_decode_entropy(padded, len(stream), C.byref(bp), out, cap)
    return out[:n].copy(), bp.value


def peek_stream(stream, pos=0):
    """(range, n, entropy_mode, prob_bits(5-bit), table_mode) of the stream header at `pos`."""
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    r, n = sz(0), sz(0)
    em, pb, tm = C.c_int(0), C.c_int(0), C.c_int(0)
    f = oracle().orc_peek_stream
    f.restype = None
    f.argtypes = [u8p, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(C.c_int), C.POINTER(C.c_int),
                  C.POINTER(C.c_int)]
    f(stream, pos, C.byref(r), C.byref(n), C.byref(em), C.byref(pb), C.byref(tm))
    return r.value, n.value, em.value, pb.value, tm.value


def synth_rgb(w, h, seed):
    a = np.zeros(w * h * 3, np.uint8)
    oracle().orc_synth_rgb(a, w, h, seed)
    return a


def synth_symbols(n, seed):
    a = np.zeros(n, np.uint8)
    oracle().orc_synth_symbols(a, n, seed)
    return a


def orc_find_lz_rgb(rgb, width, distance, bonus):
    """lz.hpp:6 -> (lz bytes, nuke map, [since_last, length-4, back%256, back/256] raw side streams)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    size = rgb.size
    cap = size // 9 + 1
    out = np.zeros(4 * (cap * 2 + 2048), np.uint8)
    nuke = np.zeros(size // 3, np.uint8)
    side = [np.zeros(cap, np.uint8) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in side])
    cnt = (sz * 4)()
    n = oracle().orc_find_lz_rgb(rgb, size, width, distance, bonus, out, nuke, ptrs, cnt)
    return out[:n].copy(), nuke, [side[k][:cnt[k]].copy() for k in range(4)]


def ref_find_lz_rgb(rgb, width, height, distance, bonus):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    size = rgb.size
    out = np.zeros(4 * ((size // 9 + 1) * 2 + 2048), np.uint8)
    nuke = np.zeros(size // 3, np.uint8)
    n = ref().ref_find_lz_rgb(rgb, size, width, height, out, nuke, distance, bonus)
    return out[:n].copy(), nuke


def orc_lz_params(rgb, mode):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    d, b = C.c_int(), C.c_int()
    oracle().orc_lz_params(rgb, rgb.size, mode, C.byref(d), C.byref(b))
    return d.value, b.value


def lz_test_image(rng, w, h, kind):
    """Images that actually contain LZ matches (the §8(d) generator has none)."""
    if kind == "flat":          # large uniform areas with a few marks
        img = np.full((h, w, 3), 200, np.uint8)
        for _ in range(12):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            img[y:y + int(rng.integers(1, 9)), x:x + int(rng.integers(1, 30))] = rng.integers(0, 256, 3)
    elif kind == "pattern":     # periodic texture: many equally long candidates at different distances
        tile = rng.integers(0, 256, (int(rng.integers(2, 7)), int(rng.integers(3, 23)), 3)).astype(np.uint8)
        img = np.tile(tile, (h // tile.shape[0] + 1, w // tile.shape[1] + 1, 1))[:h, :w].copy()
        noise = rng.random((h, w)) < 0.01
        img[noise] = rng.integers(0, 256, (int(noise.sum()), 3))
    elif kind == "rows":        # rows repeated further up: only the width-multiple distances find them
        base = rng.integers(0, 256, (4, w, 3)).astype(np.uint8)
        img = base[rng.integers(0, 4, h)].copy()
        noise = rng.random((h, w)) < 0.02
        img[noise] = rng.integers(0, 256, (int(noise.sum()), 3))
    elif kind == "few":         # a handful of colours: break-even bonus territory
        pal = rng.integers(0, 256, (int(rng.integers(2, 20)), 3)).astype(np.uint8)
        img = pal[(rng.integers(0, len(pal), (h, w)) * (rng.random((h, w)) < 0.3)).astype(np.int64)]
    else:                       # smooth synthetic photo: (almost) no matches
        img = synth_rgb(w, h, int(rng.integers(1, 1 << 30))).reshape(h, w, 3)
    return np.ascontiguousarray(img, dtype=np.uint8)


def photo_with_repeats(rng, w, h, seed):
    """A photographic image (> 256 colours, so encode_tile stays in sub-green mode) with runs repeated a short
    distance to the left and smeared runs, so that find_lz_rgb finds matches inside its 64-pixel -s0 window."""
    img = synth_rgb(w, h, seed).reshape(h, w, 3).copy()
    for _ in range(max(8, w * h // 2000)):
        n, d = int(rng.integers(4, 50)), int(rng.integers(1, 60))
        y, x = int(rng.integers(0, h)), int(rng.integers(d, max(d + 1, w - n)))
        n = min(n, w - x)
        for k in range(n):                      # forward copy: overlapping runs behave like LZ copies
            img[y, x + k] = img[y, x + k - d]
    return img


def orc_subtract_green(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    n = rgb.size // 3
    g, rg, bg = (np.zeros(n, np.uint16) for _ in range(3))
    oracle().orc_subtract_green(rgb, rgb.size, g, rg, bg)
    return g, rg, bg


def varint_bytes(v):
    out = np.zeros(4, np.uint8)
    n = oracle().orc_write_varint(out, 0, v)
    return out[:n].tobytes()


def orc_channel_picker(src, total, target):
    src = np.ascontiguousarray(src, dtype=np.uint8).ravel()
    out = np.zeros(src.size // total, np.uint16)
    oracle().orc_channel_picker(src, src.size, total, target, out)
    return out


def assemble_tile(lz, colour_mode, chans):
    """encode_tile's emit rules for three channels (choh.cpp:112-116, 328-363)."""
    return (bytes([0, 0, colour_mode]) + bytes(lz) + bytes([0b00100100]) + varint_bytes(len(chans[0])) +
            varint_bytes(len(chans[1])) + b"".join(chans))


def pick_colour_mode(sub_green, alt_rb):
    """choh.cpp:257-327 for a tile that is neither grey nor palettable: sub-green (128) unless the plain R, B
    planes (only tried at cruncher mode > 2) make the three channels smaller -> (colour mode, channels)."""
    if alt_rb is not None:
        rgb_size = len(alt_rb[0]) + len(sub_green[0]) + len(alt_rb[1])      # :289
        if rgb_size < sum(len(c) for c in sub_green):                       # :309 (the LZ bytes are on both sides)
            return 2, [sub_green[0], alt_rb[0], alt_rb[1]]
    return 128, list(sub_green)


def orc_encode_tile_subgreen(tile, mode=0):
    """encode_tile (choh.cpp:104-382) for a photographic tile (not grey, > 256 colours): header, LZ record,
    colour mode competition, channel-order byte, size varints, three channel payloads.
    tile: (h, w, 3) u8.  Returns (bytes, nuke map)."""
    tile = np.ascontiguousarray(tile, dtype=np.uint8)
    h, w = tile.shape[:2]
    distance, bonus = orc_lz_params(tile, mode)
    lz, nuke, _ = orc_find_lz_rgb(tile, w, distance, bonus)
    g, rg, bg = orc_subtract_green(tile)
    ch = [orc_layer_encode(p, w, h, d, mode, nuke)[0].tobytes() for p, d in ((g, 8), (rg, 9), (bg, 9))]
    alt = None
    if mode > 2:                                                            # :263-290
        alt = [orc_layer_encode(orc_channel_picker(tile, 3, c), w, h, 8, mode, nuke)[0].tobytes() for c in (0, 2)]
    colour_mode, chans = pick_colour_mode(ch, alt)
    return assemble_tile(lz.tobytes(), colour_mode, chans), nuke


def orc_layer_encode(plane, w, h, depth, mode, nuke=None):
    plane = np.ascontiguousarray(plane, dtype=np.uint16)
    if nuke is None:
        nuke = np.zeros(plane.size, np.uint8)
    out = np.zeros(plane.size * 4 + 4096, np.uint8)
    trace = (C.c_int * 4)()
    n = oracle().orc_layer_encode(plane, plane.size, w, h, depth, mode, nuke, out, trace)
    return out[:n].copy(), list(trace)


def ref_layer_encode(plane, w, h, depth, mode, nuke=None):
    plane = np.ascontiguousarray(plane, dtype=np.uint16)
    if nuke is None:
        nuke = np.zeros(plane.size, np.uint8)
    out = np.zeros(plane.size * 4 + 4096, np.uint8)
    n = ref().ref_layer_encode(plane, plane.size, w, h, depth, mode, nuke, out)
    return out[:n].copy()
