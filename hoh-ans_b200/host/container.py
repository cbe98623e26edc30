"""Host side of the .hoh container (choh.cpp:436-506): file header, tile grid bytes, tile-size varints and the
concatenation of the tiles' bytes.  Pure byte shuffling — the tiles themselves come from the device
(`hoh_encode_images`).  Mirrors what the reference's `main` does around `encode_tile`, including the defect that
an untiled image's single tile is never written."""


def write_varint(value):
    """varint.hpp:29-45: big-endian 7-bit groups, at most three (values >= 2^21 emit nothing, D5)."""
    if value < (1 << 7):
        return bytes([value])
    if value < (1 << 14):
        return bytes([0x80 | (value >> 7), value & 0x7f])
    if value < (1 << 21):
        return bytes([0x80 | (value >> 14), 0x80 | ((value >> 7) & 0x7f), value & 0x7f])
    return b""


def file_header(width, height):
    """choh.cpp:436-451: magic, colour format 2 (RGB), 8 bits, width-1 and height-1 as varints."""
    return bytes([153, 72, 79, 72, 2, 8]) + write_varint(width - 1) + write_varint(height - 1)


def assemble_file(width, height, x_tiles, y_tiles, tiles, flags=None):
    """choh.cpp:454-506.  tiles: the encode_tile bytes of every tile in raster order.  Returns the file bytes
    and the size `choh` prints.  flags: the `flags` column of hoh_encode_images' tile records; a tile flagged
    HOH_TILE_GREY / HOH_TILE_PALETTE was emitted in subtract-green mode where the reference would have taken its
    greyscale / indexed branch, so the file would be valid but NOT what choh writes: refused unless flags is None."""
    if flags is not None and any(int(f) for f in flags):
        raise ValueError("tiles flagged HOH_TILE_GREY / HOH_TILE_PALETTE: the reference codes them in a colour mode "
                         "that stays on the host (choh.cpp:180-213, 298-307)")
    out = bytearray(file_header(width, height))
    if (width >= 512 or height >= 512) and width >= 256 and height >= 256:
        assert len(tiles) == x_tiles * y_tiles
        out += bytes([x_tiles - 1, y_tiles - 1])
        for t in tiles[:-1]:                      # :492-494 every size but the last
            out += write_varint(len(t))
        for t in tiles:
            out += t
        return bytes(out), len(out)
    return bytes(out), len(out) + len(tiles[0])   # :508-519 the tile is sized but never written (D1)
