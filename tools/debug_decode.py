"""Development aid: walk one tile's bytes with the oracle's stream parser and print every entropy stream."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_lib, oracle_lib as ol

g = gpu_lib.gpu()
w, h, mode, n = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(17 * mode + w)
imgs = [ol.photo_with_repeats(rng, w, h, 700 + mode + i) for i in range(n)]
rgb = np.concatenate([i.ravel() for i in imgs])
tiles, rec = g.encode_images(rgb, n, w, h, mode, 24)
back, st = g.decode_images(tiles, n, w, h)
print("status", st, "rec", rec[0])
t = np.frombuffer(tiles[0], np.uint8)
print("head", tiles[0][:8].hex())
pos = 4
last_sym = None
def stream(label):
    global pos, last_sym
    sym, end, s = ol.orc_decode_entropy(t, pos, flags=7, cap=1 << 20)
    last_sym = sym
    got = g.decode_entropy_batch(np.concatenate([t, np.zeros(64, np.uint8)]), [pos], [1 << 17], flags=7)[0]
    print(label, "at", pos, "bytes", t[pos:pos + 6].tobytes().hex(), "oracle n", len(sym), "end", end, "st", s,
          "| gpu n", len(got[0]), "end", got[1], "st", got[2], "same", np.array_equal(got[0], sym))
    pos = end
lzs = []
for k in range(4):
    if k == 3 and not (t[pos] == 0x81 and t[pos + 1] == 0x7f):
        break
    stream(f"lz{k}")
    lzs.append(last_sym)
geo = g.tile_geometry(w, h)
tw, th = geo.tile_w, geo.tile_h
backref = np.zeros(tw * th, np.uint16)
p_, grp, i = 0, 0, 0
while i < len(lzs[0]):
    cnt = 0
    while i < len(lzs[0]) and lzs[0][i] == 255:
        cnt += 255; i += 1
    if i >= len(lzs[0]): break
    cnt += int(lzs[0][i]); i += 1
    p_ += cnt
    ln = int(lzs[1][grp]) + 4
    bk = int(lzs[2][grp]) + (int(lzs[3][grp]) << 8 if len(lzs) > 3 else 0)
    grp += 1
    backref[p_:p_ + ln] = bk
    p_ += ln
tile0 = np.ascontiguousarray(imgs[0][:th, :tw])
true_planes = ol.orc_subtract_green(tile0)
back_tile = back[: w * h * 3].reshape(h, w, 3)[:th, :tw]
print("tile0 rgb mismatches", int(np.count_nonzero(back_tile != tile0)), "covered px", int(np.count_nonzero(backref)))
print("order byte", hex(t[pos])); pos += 1
def varint():
    global pos
    b0 = int(t[pos]); pos += 1
    if not b0 & 0x80: return b0
    b1 = int(t[pos]); pos += 1
    if not b1 & 0x80: return ((b0 & 0x7f) << 7) + b1
    b2 = int(t[pos]); pos += 1
    return ((b0 & 0x7f) << 14) + ((b1 & 0x7f) << 7) + b2
s1, s2 = varint(), varint()
chan = [pos, pos + s1, pos + s1 + s2, len(t)]
print("channels", chan)
for c in range(3):
    pos = chan[c]
    print("chan", c, "hdr", t[pos:pos + 12].tobytes().hex())
    gx, gy = t[pos + 1] + 1, t[pos + 2] + 1
    pos += 3
    depth = 8 if c == 0 else 9
    if gx == 1 and gy == 1:
        pos += 2
        tmap = None
    else:
        cnt = int(t[pos]); masks = [(int(t[pos + 1 + 2 * m]) << 8) | int(t[pos + 2 + 2 * m]) for m in range(cnt)]
        pos += 1 + 2 * cnt
        stream(f"  idx c{c}")
        tmap = np.array([masks[int(v)] for v in last_sym], np.uint16)
    stream(f"  main c{c}")
    resid = np.zeros(tw * th, np.uint16); resid[:len(last_sym)] = last_sym
    outp = np.zeros(tw * th, np.uint16)
    if tmap is None:
        ol.oracle().orc_unpredict_fastpath(resid, tw, th, depth, backref.ctypes.data, outp)
    else:
        ol.oracle().orc_unpredict_all(resid, tw, th, depth, int(gx), int(gy), tmap, backref.ctypes.data, outp)
    bad = np.flatnonzero(outp != true_planes[c])
    gpu_plane = [back_tile[..., 1].ravel().astype(np.int32), None, None]
    print("   oracle unpredict mismatches", len(bad), "first", bad[:3], "n resid", len(last_sym), "uncovered", int(np.count_nonzero(backref == 0)))
    if c == 0:
        badg = np.flatnonzero(gpu_plane[0] != true_planes[0])
        print("   gpu green mismatches", len(badg), "first", badg[:5], [divmod(int(v), tw) for v in badg[:3]])
    print("   channel end", chan[c + 1], "parsed end", pos)

print("---- per tile")
per_img = geo.tiles_per_image
for ti, tb in enumerate(tiles):
    img_i, tl = divmod(ti, per_img)
    x0, y0 = (tl % geo.x_tiles) * tw, (tl // geo.x_tiles) * th
    want = imgs[img_i][y0:y0 + th, x0:x0 + tw]
    got = back[img_i * w * h * 3:(img_i + 1) * w * h * 3].reshape(h, w, 3)[y0:y0 + th, x0:x0 + tw]
    bad = [int(np.count_nonzero(got[..., c] != want[..., c])) for c in range(3)]
    tt = np.frombuffer(tb, np.uint8)
    print("tile", ti, "mismatch per channel RGB", bad, "rec", rec[ti])
    if sum(bad):
        pos = 4
        t = tt
        lzs = []
        for k in range(4):
            if k == 3 and not (t[pos] == 0x81 and t[pos + 1] == 0x7f):
                break
            stream(f"  lz{k}")
        pos += 1
        s1, s2 = varint(), varint()
        chan = [pos, pos + s1, pos + s1 + s2, len(t)]
        for c in range(3):
            pos = chan[c]
            print("  chan", c, "hdr", t[pos:pos + 12].tobytes().hex())
            gx, gy = t[pos + 1] + 1, t[pos + 2] + 1
            pos += 3
            if gx == 1 and gy == 1:
                pos += 2
            else:
                cnt = int(t[pos]); pos += 1 + 2 * cnt
                stream(f"    idx c{c}")
            stream(f"    main c{c}")
            print("     channel end", chan[c + 1], "parsed end", pos)
        first = np.flatnonzero((got != want).any(axis=2).ravel())[:5]
        print("  first bad pixels", [divmod(int(v), tw) for v in first])
