"""Test helper: a minimal PNG reader used to check the product's PNG writer."""
import numpy as np


def decode_png(path):
    """Minimal PNG reader (8-bit RGB, filter 0) that checks every chunk CRC: (width, height, H x W x 3 uint8)."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr, names = 8, b"", None, []
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + body) & 0xffffffff, typ
        names.append(typ)
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    assert names[0] == b"IHDR" and names[-1] == b"IEND"
    w, h, depth, colour, comp, filt, interlace = hdr
    assert (depth, colour, comp, filt, interlace) == (8, 2, 0, 0, 0)      # ColorType::Rgb, BitDepth::Eight (raytrace.rs:1464-1465)
    lines = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert not lines[:, 0].any()
    return w, h, lines[:, 1:].reshape(h, w, 3)
