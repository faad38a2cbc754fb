"""ORACLE (test infrastructure): baseline JPEG decoding restated on the CPU, byte-exact against what the reference gets
from Pillow (`Image.open(path)` ... `.convert('RGB')`, utils/image_loading.py:90-106; Pillow decodes with libjpeg-turbo:
JDCT_ISLOW, fancy upsampling, integer YCbCr tables).

libjpeg(-turbo) is a third-party dependency of the reference (through Pillow, `pillow>=10.0.0`, requirements.txt) and
is not vendored in /root/reference; its published algorithms are restated here:
  * entropy decoding         ITU T.81 Annex F.2 (sequential Huffman, restart intervals, byte stuffing)
  * dequantise + inverse DCT jidctint.c `jpeg_idct_islow` (CONST_BITS 13, PASS1_BITS 2, the 10-bit wrap of the range-limit table)
  * chroma upsampling        jdsample.c `h2v1_fancy_upsample` / `h2v2_fancy_upsample` (triangle filter, +8 / +7 and +1 / +2 biases,
                             edge rows and columns replicated, jdmainct.c context rows)
  * colour conversion        jdcolor.c `ycc_rgb_convert` (16-bit fixed-point tables, ONE_HALF folded into the Cb->G table)
Pinned by tests/test_oracle_jpeg.py against Pillow's own decoder on 4:4:4 / 4:2:2 / 4:2:0 / grayscale streams of odd and
even sizes, with and without restart markers.  Pure-Python entropy loop: small fixtures only.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], dtype=np.int64)      # zigzag position -> natural (row-major) index


class JpegHeader:
    """Everything the scan needs: frame size, per-component sampling / table ids, tables, restart interval, scan bytes."""

    def __init__(self):
        self.width = self.height = 0
        self.components = []          # dicts: id, h, v, tq, td, ta
        self.qtables = {}             # id -> int array [64] in natural order
        self.huff = {}                # (class, id) -> (bits[17], values)
        self.restart_interval = 0
        self.scan_offset = 0          # first byte of the entropy-coded segment
        self.scan_end = 0             # offset of the EOI marker (or len(data))
        self.progressive = False


def parse_header(data: bytes) -> JpegHeader:
    h = JpegHeader()
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG stream (no SOI)")
    pos = 2
    n = len(data)
    while pos < n:
        if data[pos] != 0xFF:
            raise ValueError(f"marker expected at byte {pos}")
        while data[pos] == 0xFF:
            pos += 1
        marker = data[pos]
        pos += 1
        if marker in (0xD8, 0x01) or 0xD0 <= marker <= 0xD7:
            continue
        if marker == 0xD9:
            break
        seglen = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + seglen]
        if marker == 0xDB:                                   # DQT
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    vals = list(seg[i:i + 64])
                    i += 64
                q = np.zeros(64, np.int64)
                q[ZIGZAG] = vals
                h.qtables[tq] = q
        elif marker in (0xC0, 0xC1, 0xC2):                   # SOF0 / SOF1 / SOF2
            h.progressive = marker == 0xC2
            if seg[0] != 8:
                raise ValueError("only 8-bit JPEG is supported")
            h.height, h.width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            for c in range(seg[5]):
                cid, hv, tq = seg[6 + 3 * c:9 + 3 * c]
                h.components.append({"id": cid, "h": hv >> 4, "v": hv & 15, "tq": tq})
        elif 0xC3 <= marker <= 0xCF and marker not in (0xC4, 0xC8, 0xCC):
            raise ValueError(f"unsupported JPEG process (SOF marker 0x{marker:02x})")
        elif marker == 0xC4:                                 # DHT
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                bits = [0] + list(seg[i + 1:i + 17])
                nv = sum(bits)
                h.huff[(tc, th)] = (bits, list(seg[i + 17:i + 17 + nv]))
                i += 17 + nv
        elif marker == 0xDD:                                 # DRI
            h.restart_interval = (seg[0] << 8) | seg[1]
        elif marker == 0xDA:                                 # SOS
            ns = seg[0]
            for k in range(ns):
                cs, tdta = seg[1 + 2 * k], seg[2 + 2 * k]
                for comp in h.components:
                    if comp["id"] == cs:
                        comp["td"], comp["ta"] = tdta >> 4, tdta & 15
            if ns != len(h.components):
                raise ValueError("non-interleaved multi-scan JPEG is not supported")
            h.scan_offset = pos + seglen
            # the entropy-coded segment runs to the next marker that is neither a stuffed zero nor RSTn
            j = h.scan_offset
            while j < n - 1:
                if data[j] == 0xFF and data[j + 1] != 0x00 and not (0xD0 <= data[j + 1] <= 0xD7):
                    break
                j += 1
            h.scan_end = j
            return h
        pos += seglen
    raise ValueError("no SOS marker")


def _decode_tables(bits, values):
    """Annex C: code lengths -> {(length, code): symbol}."""
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(bits[length]):
            table[(length, code)] = values[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data, pos, end):
        self.d, self.p, self.end = data, pos, end
        self.acc = 0
        self.n = 0

    def _fill(self):
        while self.n <= 24:
            if self.p < self.end:
                b = self.d[self.p]
                if b == 0xFF:
                    nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nxt == 0x00:
                        self.p += 2
                    else:                                     # a marker: feed zeros, do not advance (F.2.2.5)
                        b = 0
                else:
                    self.p += 1
            else:
                b = 0
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFFFF
            self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        self._fill()
        v = (self.acc >> (self.n - k)) & ((1 << k) - 1)
        self.n -= k
        return v

    def huff(self, table):
        self._fill()
        code = 0
        for length in range(1, 17):
            code = (code << 1) | ((self.acc >> (self.n - length)) & 1)
            sym = table.get((length, code))
            if sym is not None:
                self.n -= length
                return sym
        raise ValueError("bad Huffman code")

    def restart(self):
        """Skip to just after the next RSTn marker and drop buffered bits."""
        self.n = 0
        self.acc = 0
        p = self.p
        while p < self.end - 1 and not (self.d[p] == 0xFF and 0xD0 <= self.d[p + 1] <= 0xD7):
            p += 1
        self.p = p + 2


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def geometry(h: JpegHeader):
    hmax = max(c["h"] for c in h.components)
    vmax = max(c["v"] for c in h.components)
    mcux = -(-h.width // (8 * hmax))
    mcuy = -(-h.height // (8 * vmax))
    return hmax, vmax, mcux, mcuy


def decode_coefficients(data: bytes, h: JpegHeader):
    """Quantised coefficients per component: int16 [block_rows][block_cols][64] in natural (row-major) order."""
    if h.progressive:
        raise ValueError("progressive JPEG is not supported")
    hmax, vmax, mcux, mcuy = geometry(h)
    coefs = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int16) for c in h.components]
    dct = [_decode_tables(*h.huff[(0, c["td"])]) for c in h.components]
    act = [_decode_tables(*h.huff[(1, c["ta"])]) for c in h.components]
    br = _Bits(data, h.scan_offset, h.scan_end)
    pred = [0] * len(h.components)
    ri = h.restart_interval
    for m in range(mcux * mcuy):
        if ri and m and m % ri == 0:
            br.restart()
            pred = [0] * len(h.components)
        my, mx = divmod(m, mcux)
        for ci, c in enumerate(h.components):
            for by in range(c["v"]):
                for bx in range(c["h"]):
                    blk = coefs[ci][my * c["v"] + by, mx * c["h"] + bx]
                    t = br.huff(dct[ci])
                    pred[ci] += _extend(br.get(t), t) if t else 0
                    blk[0] = pred[ci]
                    k = 1
                    while k < 64:
                        rs = br.huff(act[ci])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k]] = _extend(br.get(s), s)
                        k += 1
    return coefs


# ---- jidctint.c: jpeg_idct_islow ---------------------------------------------------------------------------------------
CONST_BITS, PASS1_BITS = 13, 2
FIX_0_298631336, FIX_0_390180644, FIX_0_541196100, FIX_0_765366865 = 2446, 3196, 4433, 6270
FIX_0_899976223, FIX_1_175875602, FIX_1_501321110, FIX_1_847759065 = 7373, 9633, 12299, 15137
FIX_1_961570560, FIX_2_053119869, FIX_2_562915447, FIX_3_072711026 = 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(d, shift_in_bits, descale_bits):
    """One pass over the LAST axis of d (int64 [..., 8]); the even part enters shifted by CONST_BITS."""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * FIX_0_541196100
    tmp2 = z1 + z3 * (-FIX_1_847759065)
    tmp3 = z1 + z2 * FIX_0_765366865
    z2, z3 = d[..., 0], d[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * FIX_1_175875602
    t0 = t0 * FIX_0_298631336
    t1 = t1 * FIX_2_053119869
    t2 = t2 * FIX_3_072711026
    t3 = t3 * FIX_1_501321110
    z1 = z1 * (-FIX_0_899976223)
    z2 = z2 * (-FIX_2_562915447)
    z3 = z3 * (-FIX_1_961570560) + z5
    z4 = z4 * (-FIX_0_390180644) + z5
    t0 = t0 + z1 + z3
    t1 = t1 + z2 + z4
    t2 = t2 + z2 + z3
    t3 = t3 + z1 + z4
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], axis=-1)
    return _descale(out, descale_bits)


def idct_islow(coefs: np.ndarray, qtable: np.ndarray) -> np.ndarray:
    """coefs int16 [..., 64] (natural order), qtable [64] -> uint8 samples [..., 8, 8]."""
    d = coefs.astype(np.int64) * qtable.astype(np.int64)
    d = d.reshape(d.shape[:-1] + (8, 8))
    # pass 1: columns (the input index runs down a column), results scaled up by 2^PASS1_BITS
    ws = _idct_1d(np.swapaxes(d, -1, -2), 0, CONST_BITS - PASS1_BITS)
    ws = np.swapaxes(ws, -1, -2)
    # pass 2: rows; descale by CONST_BITS + PASS1_BITS + 3, then the range-limit table with its 10-bit wrap
    out = _idct_1d(ws, 0, CONST_BITS + PASS1_BITS + 3)
    v10 = ((out & 1023) ^ 512) - 512
    return np.clip(v10 + 128, 0, 255).astype(np.uint8)


def component_planes(coefs, h: JpegHeader):
    """Per component the decoded sample plane, cropped to its downsampled size (ceil(size * samp / max_samp))."""
    hmax, vmax, _, _ = geometry(h)
    planes = []
    for ci, c in enumerate(h.components):
        blocks = idct_islow(coefs[ci], h.qtables[c["tq"]])             # [by][bx][8][8]
        by, bx = blocks.shape[:2]
        plane = blocks.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)
        dw = -(-h.width * c["h"] // hmax)
        dh = -(-h.height * c["v"] // vmax)
        planes.append(np.ascontiguousarray(plane[:dh, :dw]))
    return planes


# ---- jdsample.c -----------------------------------------------------------------------------------------------------------
def h2v1_fancy_upsample(p: np.ndarray) -> np.ndarray:
    """[rows][w] -> [rows][2w]: 3/4 nearer + 1/4 further, biases +1 (left sample) / +2 (right sample)."""
    x = p.astype(np.int64)
    rows, w = x.shape
    out = np.empty((rows, 2 * w), np.int64)
    if w == 1:
        out[:, 0] = out[:, 1] = x[:, 0]
        return out.astype(np.uint8)
    left = np.concatenate([x[:, :1], x[:, :-1]], axis=1)
    right = np.concatenate([x[:, 1:], x[:, -1:]], axis=1)
    out[:, 0::2] = (x * 3 + left + 1) >> 2
    out[:, 1::2] = (x * 3 + right + 2) >> 2
    out[:, 0] = x[:, 0]
    out[:, -1] = x[:, -1]
    return out.astype(np.uint8)


def h2v2_fancy_upsample(p: np.ndarray) -> np.ndarray:
    """[rows][w] -> [2 rows][2 w]: triangle filter in both directions, biases +8 / +7; the row above the first and
    below the last row are copies of those rows (jdmainct.c), the first / last column are special-cased."""
    x = p.astype(np.int64)
    rows, w = x.shape
    up = np.concatenate([x[:1], x[:-1]], axis=0)
    dn = np.concatenate([x[1:], x[-1:]], axis=0)
    out = np.empty((2 * rows, 2 * w), np.int64)
    for v, nb in ((0, up), (1, dn)):
        colsum = x * 3 + nb                                   # "thiscolsum" per column
        if w == 1:
            row = np.concatenate([(colsum * 4 + 8) >> 4, (colsum * 4 + 7) >> 4], axis=1)
        else:
            last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
            nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
            row = np.empty((rows, 2 * w), np.int64)
            row[:, 0::2] = (colsum * 3 + last + 8) >> 4
            row[:, 1::2] = (colsum * 3 + nxt + 7) >> 4
            row[:, 0] = (colsum[:, 0] * 4 + 8) >> 4
            row[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[v::2] = row
    return out.astype(np.uint8)


# ---- jdcolor.c ------------------------------------------------------------------------------------------------------------
SCALEBITS = 16
ONE_HALF = 1 << (SCALEBITS - 1)


def _fix(x):
    return int(x * (1 << SCALEBITS) + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + ONE_HALF) >> SCALEBITS
CB_B = (_fix(1.77200) * _X + ONE_HALF) >> SCALEBITS
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + ONE_HALF


def ycc_to_rgb(y, cb, cr):
    y = y.astype(np.int64)
    r = np.clip(y + CR_R[cr], 0, 255)
    g = np.clip(y + ((CB_G[cb] + CR_G[cr]) >> SCALEBITS), 0, 255)
    b = np.clip(y + CB_B[cb], 0, 255)
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def decode_rgb(data: bytes) -> np.ndarray:
    """Baseline JPEG bytes -> [H][W][3] uint8 RGB, as `np.asarray(Image.open(...).convert('RGB'))`."""
    h = parse_header(data)
    coefs = decode_coefficients(data, h)
    planes = component_planes(coefs, h)
    if len(planes) == 1:
        return np.repeat(planes[0][:h.height, :h.width, None], 3, axis=2)
    if len(planes) != 3:
        raise ValueError("only grayscale and YCbCr JPEG are supported")
    hmax, vmax, _, _ = geometry(h)
    full = []
    for c, p in zip(h.components, planes):
        fh, fv = hmax // c["h"], vmax // c["v"]
        if (fh, fv) == (1, 1):
            u = p
        elif (fh, fv) == (2, 1):
            u = h2v1_fancy_upsample(p)
        elif (fh, fv) == (2, 2):
            u = h2v2_fancy_upsample(p)
        else:
            raise ValueError(f"unsupported sampling ratio {fh}x{fv}")      # 4:4:0 etc.: Pillow cannot write them, nothing to pin against
        full.append(u[:h.height, :h.width])
    return ycc_to_rgb(*full)
