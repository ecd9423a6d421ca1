"""BGZF (blocked gzip) helpers for tests, tools and the synthetic WGS generator: a writer and the block walker the C++
driver mirrors (rustseq_host.cpp).  A BGZF file is a plain multi-member gzip file -- `zcat` reads it -- whose members are
<= 64 KiB, each announcing its compressed size in a 'BC' extra field, which is what lets a GPU inflate them in parallel."""
import struct
import zlib

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def compress(data, level=1, block_size=65280, eof=True):
    out = []
    for a in range(0, len(data), block_size):
        chunk = data[a:a + block_size]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(payload) + 25))
        out.append(payload)
        out.append(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    if eof:
        out.append(EOF_BLOCK)
    return b"".join(out)


def walk(buf, pos=0):
    """[(payload offset, payload length, inflated length), ...] of the whole BGZF blocks in buf[pos:], and the offset of
    the first byte not consumed (a partial block at the end stays for the next call).  Raises ValueError on non-BGZF data."""
    blocks = []
    n = len(buf)
    while pos + 18 <= n:
        if buf[pos] != 0x1F or buf[pos + 1] != 0x8B or buf[pos + 2] != 8 or not (buf[pos + 3] & 4):
            raise ValueError("not a BGZF block")
        xlen = buf[pos + 10] | (buf[pos + 11] << 8)
        if pos + 12 + xlen > n:
            break
        bsize, q = None, pos + 12
        while q + 4 <= pos + 12 + xlen:
            slen = buf[q + 2] | (buf[q + 3] << 8)
            if buf[q] == 66 and buf[q + 1] == 67 and slen == 2:
                bsize = buf[q + 4] | (buf[q + 5] << 8)
            q += 4 + slen
        if bsize is None:
            raise ValueError("gzip member without a BC field")
        total = bsize + 1
        if pos + total > n:
            break
        isize = struct.unpack_from("<I", buf, pos + total - 4)[0]
        blocks.append((pos + 12 + xlen, total - 12 - xlen - 8, isize))
        pos += total
    return blocks, pos
