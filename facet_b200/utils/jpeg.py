"""Host side of the GPU JPEG decoder: header parsing and table packing (the bytes never get decoded here).

The reference decodes files on the CPU through Pillow / libjpeg-turbo (`Image.open(path)`, `ImageOps.exif_transpose`,
`.convert('RGB')`, `cv2.cvtColor(RGB2BGR)`: utils/image_loading.py:90-106).  Here a loader hands over the file BYTES;
`parse` reads the marker segments (SOF0/SOF1, DQT, DHT, DRI, SOS, APP1 orientation) and `pack_tables` lays the
quantisation and Huffman tables out for csrc/jpeg_decode.cu.  Entropy decoding, inverse DCT, chroma upsampling and
colour conversion all run on the device (`ops.jpeg_decode`), byte-exact with Pillow's output.

Supported: baseline / extended sequential Huffman, 8 bit, 1 or 3 components in one interleaved scan, luma sampling
1x1, 2x1 or 2x2 with 1x1 chroma (4:4:4, 4:2:2, 4:2:0) — what cameras and Pillow write — with restart markers (one thread
per restart interval) or without (self-synchronising parallel decoding).  Progressive, arithmetic-coded, 12-bit, CMYK and
multi-scan files raise `UnsupportedJpeg` (the caller keeps its CPU loader for those, as the reference does for RAW files).
"""
from __future__ import annotations

import struct

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], dtype=np.int64)      # zigzag position -> natural (row-major) index

SERIAL_MCUS = 1024           # streams without restart markers up to this many MCUs are decoded by one thread; larger ones by the
                             # self-synchronising scheme of csrc/jpeg_decode.cu (kSerialMcus there)
LUT_BITS = 9
# one Huffman table on the device: uint16 lut[512] | int32 maxcode[18] | int32 valptr[17] | uint8 values[256]  (see JpegHuff in
# csrc/jpeg_decode.cu); a table set = uint16 q[4][64] + 4 DC + 4 AC tables
HUFF_BYTES = 2 * (1 << LUT_BITS) + 4 * 18 + 4 * 17 + 256 + 4          # + 4 bytes padding -> multiple of 8
TABLESET_BYTES = 2 * 4 * 64 + 8 * HUFF_BYTES


class UnsupportedJpeg(ValueError):
    pass


class JpegInfo:
    __slots__ = ("width", "height", "ncomp", "hs", "vs", "tq", "td", "ta", "restart_interval", "scan_offset", "scan_end",
                 "qtables", "huff", "orientation", "packed_tables")

    def geometry(self):
        """(hmax, vmax, mcux, mcuy)."""
        hmax, vmax = max(self.hs), max(self.vs)
        return hmax, vmax, -(-self.width // (8 * hmax)), -(-self.height // (8 * vmax))

    def geometry_key(self):
        """Streams with equal keys can share one decode launch."""
        return (self.width, self.height, self.ncomp, tuple(self.hs), tuple(self.vs), self.restart_interval)


def _exif_orientation(seg: bytes) -> int:
    """Orientation (tag 0x0112) from an APP1 Exif segment, 1 if absent — what `ImageOps.exif_transpose` reads."""
    if seg[:6] != b"Exif\x00\x00" or len(seg) < 14:
        return 1
    t = seg[6:]
    if t[:2] == b"II":
        e = "<"
    elif t[:2] == b"MM":
        e = ">"
    else:
        return 1
    try:
        off = struct.unpack(e + "I", t[4:8])[0]
        n = struct.unpack(e + "H", t[off:off + 2])[0]
        for k in range(n):
            ent = t[off + 2 + 12 * k:off + 14 + 12 * k]
            tag, typ = struct.unpack(e + "HH", ent[:4])
            if tag == 0x0112:
                val = struct.unpack(e + "H", ent[8:10])[0] if typ == 3 else struct.unpack(e + "I", ent[8:12])[0]
                return val if 1 <= val <= 8 else 1
    except (struct.error, IndexError):
        pass
    return 1


_TABLE_CACHE: dict = {}


def _tables_from_segments(dqt: bytes, dht: bytes):
    """(qtables, huff, packed table set) of the concatenated DQT / DHT segment payloads; cached, because the streams of one
    camera / encoder setting carry identical tables and parsing them costs more than everything else in `parse`."""
    key = (dqt, dht)
    hit = _TABLE_CACHE.get(key)
    if hit is not None:
        return hit
    qtables, huff = {}, {}
    i = 0
    while i < len(dqt):
        pq, tq = dqt[i] >> 4, dqt[i] & 15
        i += 1
        if pq:
            vals = np.frombuffer(dqt[i:i + 128], dtype=">u2").astype(np.int64)
            i += 128
        else:
            vals = np.frombuffer(dqt[i:i + 64], dtype=np.uint8).astype(np.int64)
            i += 64
        q = np.zeros(64, np.int64)
        q[ZIGZAG] = vals
        qtables[tq] = q
    i = 0
    while i < len(dht):
        tc, th = dht[i] >> 4, dht[i] & 15
        bits = [0] + list(dht[i + 1:i + 17])
        nv = sum(bits)
        huff[(tc, th)] = (bits, bytes(dht[i + 17:i + 17 + nv]))
        i += 17 + nv
    packed = _pack_table_set(qtables, huff)
    if len(_TABLE_CACHE) > 256:
        _TABLE_CACHE.clear()
    _TABLE_CACHE[key] = (qtables, huff, packed)
    return _TABLE_CACHE[key]


def parse(data, head_bytes: int = 1 << 16) -> JpegInfo:
    """Marker segments of one JPEG stream: bytes / bytearray, or a 1-D uint8 numpy array / memoryview (e.g. a view of a
    pinned read buffer; only the header region is copied out).  Raises UnsupportedJpeg."""
    total = len(data)
    if isinstance(data, (bytes, bytearray)):
        head = data
    else:
        head = bytes(memoryview(data)[:min(total, head_bytes)])
    if head[:2] != b"\xff\xd8":
        raise UnsupportedJpeg("not a JPEG stream (no SOI marker)")
    info = JpegInfo()
    info.restart_interval = 0
    info.orientation = 1
    comps = None
    dqt, dht = [], []
    pos, n = 2, len(head)
    while True:
        if pos + 4 > n:
            if n < total:                              # the header is longer than the window (big APPn blocks): widen it
                return parse(data, head_bytes=total)
            break
        if head[pos] != 0xFF:
            raise UnsupportedJpeg(f"marker expected at byte {pos}")
        while pos < n and head[pos] == 0xFF:
            pos += 1
        if pos >= n:
            continue
        marker = head[pos]
        pos += 1
        if marker == 0x01 or 0xD0 <= marker <= 0xD8:
            continue
        if marker == 0xD9:
            break
        seglen = (head[pos] << 8) | head[pos + 1]
        if pos + seglen > n:
            if n < total:
                return parse(data, head_bytes=total)
            raise UnsupportedJpeg("truncated marker segment")
        if marker == 0xDB:
            dqt.append(head[pos + 2:pos + seglen])
        elif marker == 0xC4:
            dht.append(head[pos + 2:pos + seglen])
        elif marker in (0xC0, 0xC1):
            seg = head[pos + 2:pos + seglen]
            if seg[0] != 8:
                raise UnsupportedJpeg("only 8-bit samples are supported")
            info.height, info.width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            comps = [(seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c]) for c in range(seg[5])]
        elif 0xC2 <= marker <= 0xCF and marker not in (0xC8, 0xCC):
            raise UnsupportedJpeg("progressive / lossless / arithmetic-coded JPEG (SOF marker 0x%02x)" % marker)
        elif marker == 0xDD:
            info.restart_interval = (head[pos + 2] << 8) | head[pos + 3]
        elif marker == 0xE1 and info.orientation == 1:
            info.orientation = _exif_orientation(head[pos + 2:pos + seglen])
        elif marker == 0xDA:
            seg = head[pos + 2:pos + seglen]
            if comps is None:
                raise UnsupportedJpeg("SOS before SOF")
            ns = seg[0]
            if ns != len(comps):
                raise UnsupportedJpeg("multi-scan (non-interleaved) JPEG")
            sel = {seg[1 + 2 * k]: seg[2 + 2 * k] for k in range(ns)}
            info.ncomp = len(comps)
            if info.ncomp not in (1, 3):
                raise UnsupportedJpeg(f"{info.ncomp}-component JPEG (only grayscale and YCbCr)")
            info.hs = [c[1] for c in comps]
            info.vs = [c[2] for c in comps]
            info.tq = [c[3] for c in comps]
            info.td = [sel[c[0]] >> 4 for c in comps]
            info.ta = [sel[c[0]] & 15 for c in comps]
            if info.ncomp == 1:
                info.hs, info.vs = [1], [1]            # a single-component scan is never interleaved: 8x8 MCUs (T.81 A.2.2)
            elif (info.hs[0], info.vs[0]) not in ((1, 1), (2, 1), (2, 2)) or info.hs[1:] != [1, 1] or info.vs[1:] != [1, 1]:
                raise UnsupportedJpeg(f"sampling factors {list(zip(info.hs, info.vs))}")
            info.qtables, info.huff, info.packed_tables = _tables_from_segments(b"".join(dqt), b"".join(dht))
            for c in range(info.ncomp):
                if (info.tq[c] not in info.qtables or info.tq[c] > 3 or (0, info.td[c]) not in info.huff
                        or (1, info.ta[c]) not in info.huff or info.td[c] > 3 or info.ta[c] > 3):
                    raise UnsupportedJpeg("a table referenced by the scan is missing")
            info.scan_offset = pos + seglen
            # the entropy-coded segment ends at the EOI marker: normally the last two bytes of the file
            tail = bytes(memoryview(data)[max(info.scan_offset, total - 4096):total]) if not isinstance(data, (bytes, bytearray)) else None
            if tail is not None:
                end = tail.rfind(b"\xff\xd9")
                if end < 0:
                    end = bytes(memoryview(data)[info.scan_offset:total]).rfind(b"\xff\xd9")
                    info.scan_end = info.scan_offset + end if end >= 0 else total
                else:
                    info.scan_end = max(info.scan_offset, total - 4096) + end
            else:
                end = data.rfind(b"\xff\xd9")
                info.scan_end = end if end >= info.scan_offset else total
            return info
        pos += seglen
    raise UnsupportedJpeg("no SOS marker")


def _pack_huff(bits, values) -> bytes:
    """Annex C code assignment -> 9-bit lookahead table + the maxcode / valptr arrays of the Annex F.2.2.3 decoder."""
    lut = np.zeros(1 << LUT_BITS, np.uint16)
    maxcode = np.full(18, -1, np.int32)
    valptr = np.zeros(17, np.int32)
    code, k = 0, 0
    for length in range(1, 17):
        valptr[length] = k - code                     # symbol index = code + valptr[length]
        for _ in range(bits[length]):
            if length <= LUT_BITS:
                lo = code << (LUT_BITS - length)
                lut[lo:lo + (1 << (LUT_BITS - length))] = (length << 8) | values[k]
            code += 1
            k += 1
        maxcode[length] = code - 1 if bits[length] else -1
        code <<= 1
    maxcode[17] = 0x7FFFFFFF
    vals = np.zeros(256, np.uint8)
    vals[:len(values)] = np.frombuffer(bytes(values), np.uint8)
    return lut.tobytes() + maxcode.tobytes() + valptr.tobytes() + vals.tobytes() + b"\x00" * 4


def _pack_table_set(qtables, huff) -> bytes:
    q = np.zeros((4, 64), np.uint16)
    for tq, tab in qtables.items():
        if tq < 4:
            q[tq] = tab
    out = [q.tobytes()]
    empty = b"\x00" * HUFF_BYTES
    for tc in (0, 1):
        for th in range(4):
            out.append(_pack_huff(*huff[(tc, th)]) if (tc, th) in huff else empty)
    blob = b"".join(out)
    assert len(blob) == TABLESET_BYTES
    return blob


def pack_tables(info: JpegInfo) -> bytes:
    """One table set (TABLESET_BYTES): q[4][64] uint16 in natural order, DC tables 0..3, AC tables 0..3."""
    return info.packed_tables


# ---- encoder side: what fb_jpeg_encode needs beside the pixels ---------------------------------------------------------------
_ZIGZAG = (0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63)
_ENC_CACHE: dict = {}


def encoder_tables(height: int, width: int, quality: int = 80):
    """(header bytes, packed tables) for `Image.save(format='JPEG', quality=quality)` of an RGB image of this size.

    The header — SOI, JFIF APP0, the two DQT and four DHT segments, SOF0, the SOS header — does not depend on the pixels, so it is
    taken from the stream Pillow (the reference's encoder, utils/image_transforms.py:47) writes for a blank image of this size; the
    quantisation and Huffman tables fb_jpeg_encode uses are parsed from that same header.  Packed layout (3328 bytes): uint16
    q8[2][64] = 8 x quantisation value in natural order (luma, chroma), uint16 code[4][256], uint8 size[4][256] for the DC luma,
    AC luma, DC chroma and AC chroma tables."""
    key = (int(height), int(width), int(quality))
    if key in _ENC_CACHE:
        return _ENC_CACHE[key]
    import io
    from PIL import Image
    buf = io.BytesIO()
    Image.new("RGB", (int(width), int(height))).save(buf, format="JPEG", quality=int(quality))
    data = buf.getvalue()
    i, q, huff = 2, {}, {}
    while True:
        if data[i] != 0xFF:
            raise ValueError("unexpected byte in the JPEG header Pillow wrote")
        marker, length = data[i + 1], (data[i + 2] << 8) | data[i + 3]
        seg = data[i + 4:i + 2 + length]
        if marker == 0xDB:
            p = 0
            while p < len(seg):
                if seg[p] >> 4:
                    raise ValueError("16-bit quantisation table")
                nat = np.zeros(64, np.uint16)
                nat[list(_ZIGZAG)] = np.frombuffer(seg[p + 1:p + 65], np.uint8)
                q[seg[p] & 15] = nat
                p += 65
        elif marker == 0xC4:
            p = 0
            while p < len(seg):
                bits = list(seg[p + 1:p + 17])
                n = sum(bits)
                huff[(seg[p] >> 4, seg[p] & 15)] = (bits, list(seg[p + 17:p + 17 + n]))
                p += 17 + n
        elif marker == 0xC0:
            # 3 components, 2x2 / 1x1 / 1x1 sampling is what the encoder kernels implement (Pillow's default at quality <= 100... )
            comps = [(seg[6 + 3 * c], seg[7 + 3 * c], seg[8 + 3 * c]) for c in range(seg[5])]
            if seg[5] != 3 or [c[1] for c in comps] != [0x22, 0x11, 0x11] or [c[2] for c in comps] != [0, 1, 1]:
                raise ValueError("Pillow did not choose YCbCr 4:2:0 with two quantisation tables for this setting")
        elif marker == 0xDA:
            header = bytes(data[:i + 2 + length])
            break
        i += 2 + length
    packed = np.zeros(3328, np.uint8)
    packed[:256].view(np.uint16)[:] = np.concatenate([q[0] * 8, q[1] * 8]).astype(np.uint16)
    code = packed[256:256 + 2048].view(np.uint16).reshape(4, 256)
    size = packed[256 + 2048:].reshape(4, 256)
    for t, key2 in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        bits, vals = huff[key2]
        c, k = 0, 0
        for length in range(1, 17):
            for _ in range(bits[length - 1]):
                code[t, vals[k]] = c
                size[t, vals[k]] = length
                c += 1
                k += 1
            c <<= 1
    _ENC_CACHE[key] = (header, packed.tobytes())
    return _ENC_CACHE[key]
