"""TEST INFRASTRUCTURE (oracle): the BGZF layer as htslib reads it (SAM spec 4.1; htslib bgzf.c bgzf_read_block ->
inflate_block, which rust-htslib's bam::IndexedReader sits on: reference src/main.rs:1745-1757).  Every member is an
independent raw DEFLATE stream; zlib (the library htslib itself links) is the checker for np2_bgzf_inflate and for the
host/device decoder in nextpolish2_b200/csrc/np2_inflate.cuh.  Only tests/ and bench.py may import this."""
import struct
import zlib


def members(buf):
    """-> list of (payload_off, payload_len, isize, crc32) in file order."""
    buf = bytes(buf)
    out, o = [], 0
    while o < len(buf):
        if buf[o:o + 3] != b"\x1f\x8b\x08" or not buf[o + 3] & 4:
            raise ValueError("not a BGZF member at byte %d" % o)
        xlen = struct.unpack_from("<H", buf, o + 10)[0]
        x, bsize = o + 12, None
        while x < o + 12 + xlen:
            si1, si2, slen = buf[x], buf[x + 1], struct.unpack_from("<H", buf, x + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", buf, x + 4)[0]
            x += 4 + slen
        total = bsize + 1
        crc, isize = struct.unpack_from("<II", buf, o + total - 8)
        out.append((o + 12 + xlen, total - 12 - xlen - 8, isize, crc))
        o += total
    return out


def inflate_member(buf, payload_off, payload_len, isize=None, crc=None):
    data = zlib.decompress(bytes(buf[payload_off:payload_off + payload_len]), -15)
    if isize is not None and len(data) != isize:
        raise ValueError("ISIZE mismatch")
    if crc is not None and zlib.crc32(data) != crc:
        raise ValueError("CRC32 mismatch")
    return data


def inflate_all(buf):
    return b"".join(inflate_member(buf, *m) for m in members(buf))


def make_member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem_level=8):
    """One BGZF member holding `data` (<= 64 KiB)."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    payload = co.compress(bytes(data)) + co.flush()
    if len(payload) + 25 > 65535:
        raise ValueError("member would exceed 64 KiB (BSIZE is 16 bits): use <= 65280 bytes of incompressible data")
    head = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(payload) + 25)
    return head + payload + struct.pack("<II", zlib.crc32(bytes(data)), len(data))
