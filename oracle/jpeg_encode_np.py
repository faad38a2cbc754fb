"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the JPEG encoder behind the reference's thumbnails.

`generate_photo_thumbnail` (utils/image_transforms.py:32-50, called at processing/scorer.py:1681-1686) ends in
`thumb.save(buf, format='JPEG', quality=80)`: Pillow driving libjpeg(-turbo) at its defaults (YCbCr 4:2:0, standard Huffman
tables, no restart markers).  libjpeg is third-party to the reference and its source is not under /root/reference; this module
restates the published baseline encoder (rgb_ycc_convert, h2v2_downsample with edge replication, jpeg_fdct_islow, round-half-up
quantisation, dummy blocks, encode_one_block with 0xFF stuffing and one-bit padding) and is PINNED against the installed Pillow
itself: `encode(rgb, template)` reproduces Pillow's byte stream exactly on photo-like and noise images from 1x1 to 640x480
(tests/test_oracle_jpeg_encode.py).  The header bytes (SOI .. SOS) are taken from `template`, a stream Pillow wrote for the same
size and quality.  Only tests may import this module; the product path is facet_b200/csrc/jpeg_encode.cu.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])

def FIX(x): return int(x * 65536 + 0.5)

def rgb_to_ycc(rgb):
    r = rgb[..., 0].astype(np.int64); g = rgb[..., 1].astype(np.int64); b = rgb[..., 2].astype(np.int64)
    half = 1 << 15; off = 128 << 16
    y = (FIX(0.29900) * r + FIX(0.58700) * g + FIX(0.11400) * b + half) >> 16
    cb = (-FIX(0.16874) * r - FIX(0.33126) * g + FIX(0.50000) * b + off + half - 1) >> 16
    cr = (FIX(0.50000) * r - FIX(0.41869) * g - FIX(0.08131) * b + off + half - 1) >> 16
    return y.astype(np.int64), cb.astype(np.int64), cr.astype(np.int64)

def fdct_islow(blocks):
    """blocks: [..., 8, 8] int64 samples (level-shifted) -> coefficients scaled by 8 (jfdctint.c)."""
    CB, P1 = 13, 2
    F = {k: int(v * (1 << CB) + 0.5) for k, v in dict(a=0.298631336, b=0.390180644, c=0.541196100, d=0.765366865, e=0.899976223,
                                                   f=1.175875602, g=1.501321110, h=1.847759065, i=1.961570560, j=2.053119869,
                                                   k=2.562915447, l=3.072711026).items()}
    def desc(x, n): return (x + (1 << (n - 1))) >> n
    d = blocks.astype(np.int64).copy()
    # pass 1: rows
    t0 = d[..., 0] + d[..., 7]; t7 = d[..., 0] - d[..., 7]
    t1 = d[..., 1] + d[..., 6]; t6 = d[..., 1] - d[..., 6]
    t2 = d[..., 2] + d[..., 5]; t5 = d[..., 2] - d[..., 5]
    t3 = d[..., 3] + d[..., 4]; t4 = d[..., 3] - d[..., 4]
    t10 = t0 + t3; t13 = t0 - t3; t11 = t1 + t2; t12 = t1 - t2
    o = np.empty_like(d)
    o[..., 0] = (t10 + t11) << P1
    o[..., 4] = (t10 - t11) << P1
    z1 = (t12 + t13) * F['c']
    o[..., 2] = desc(z1 + t13 * F['d'], CB - P1)
    o[..., 6] = desc(z1 + t12 * (-F['h']), CB - P1)
    z1 = t4 + t7; z2 = t5 + t6; z3 = t4 + t6; z4 = t5 + t7; z5 = (z3 + z4) * F['f']
    t4 = t4 * F['a']; t5 = t5 * F['j']; t6 = t6 * F['l']; t7 = t7 * F['g']
    z1 = z1 * (-F['e']); z2 = z2 * (-F['k']); z3 = z3 * (-F['i']) + z5; z4 = z4 * (-F['b']) + z5
    o[..., 7] = desc(t4 + z1 + z3, CB - P1)
    o[..., 5] = desc(t5 + z2 + z4, CB - P1)
    o[..., 3] = desc(t6 + z2 + z3, CB - P1)
    o[..., 1] = desc(t7 + z1 + z4, CB - P1)
    # pass 2: columns
    d = o
    t0 = d[..., 0, :] + d[..., 7, :]; t7 = d[..., 0, :] - d[..., 7, :]
    t1 = d[..., 1, :] + d[..., 6, :]; t6 = d[..., 1, :] - d[..., 6, :]
    t2 = d[..., 2, :] + d[..., 5, :]; t5 = d[..., 2, :] - d[..., 5, :]
    t3 = d[..., 3, :] + d[..., 4, :]; t4 = d[..., 3, :] - d[..., 4, :]
    t10 = t0 + t3; t13 = t0 - t3; t11 = t1 + t2; t12 = t1 - t2
    o = np.empty_like(d)
    o[..., 0, :] = desc(t10 + t11, P1)
    o[..., 4, :] = desc(t10 - t11, P1)
    z1 = (t12 + t13) * F['c']
    o[..., 2, :] = desc(z1 + t13 * F['d'], CB + P1)
    o[..., 6, :] = desc(z1 + t12 * (-F['h']), CB + P1)
    z1 = t4 + t7; z2 = t5 + t6; z3 = t4 + t6; z4 = t5 + t7; z5 = (z3 + z4) * F['f']
    t4 = t4 * F['a']; t5 = t5 * F['j']; t6 = t6 * F['l']; t7 = t7 * F['g']
    z1 = z1 * (-F['e']); z2 = z2 * (-F['k']); z3 = z3 * (-F['i']) + z5; z4 = z4 * (-F['b']) + z5
    o[..., 7, :] = desc(t4 + z1 + z3, CB + P1)
    o[..., 5, :] = desc(t5 + z2 + z4, CB + P1)
    o[..., 3, :] = desc(t6 + z2 + z3, CB + P1)
    o[..., 1, :] = desc(t7 + z1 + z4, CB + P1)
    return o

def parse_header(data):
    """Quantisation tables (natural order), Huffman specs and the offset of the entropy-coded data from a Pillow-written stream."""
    i = 2; q = {}; huff = {}
    while True:
        assert data[i] == 0xFF
        m = data[i + 1]; L = (data[i + 2] << 8) | data[i + 3]; seg = data[i + 4:i + 2 + L]
        if m == 0xDB:
            p = 0
            while p < len(seg):
                pq, tq = seg[p] >> 4, seg[p] & 15
                tab = np.frombuffer(seg[p + 1:p + 65], np.uint8).astype(np.int64)
                nat = np.zeros(64, np.int64); nat[ZIGZAG] = tab
                q[tq] = nat; p += 65
        elif m == 0xC4:
            p = 0
            while p < len(seg):
                tc, th = seg[p] >> 4, seg[p] & 15
                bits = list(seg[p + 1:p + 17]); n = sum(bits)
                vals = list(seg[p + 17:p + 17 + n]); huff[(tc, th)] = (bits, vals); p += 17 + n
        elif m == 0xDA:
            return q, huff, i + 2 + L
        i += 2 + L

def huff_codes(bits, vals):
    code = 0; k = 0; ehufco = {}; ehufsi = {}
    for l in range(1, 17):
        for _ in range(bits[l - 1]):
            ehufco[vals[k]] = code; ehufsi[vals[k]] = l; code += 1; k += 1
        code <<= 1
    return ehufco, ehufsi

def encode(rgb, template):
    """template: Pillow-encoded stream of an image of the same size / quality (header source).  Returns the full JPEG bytes."""
    H, W = rgb.shape[:2]
    q, huff, sos_end = parse_header(template)
    y, cb, cr = rgb_to_ycc(rgb)
    mcux, mcuy = (W + 15) // 16, (H + 15) // 16
    # component geometry (jpeg_component_info)
    yw_blocks, yh_blocks = (W + 7) // 8, (H + 7) // 8
    cw, chh = (W + 1) // 2, (H + 1) // 2
    cw_blocks, ch_blocks = (cw + 7) // 8, (chh + 7) // 8
    def pad_edge(p, hh, ww):
        ph, pw = p.shape
        return np.pad(p, ((0, hh - ph), (0, ww - pw)), mode="edge")
    # luma: right edge replicated to width_in_blocks * 8, bottom to height_in_blocks... rows beyond the image replicate the last row
    Y = pad_edge(y, yh_blocks * 8, yw_blocks * 8)
    # chroma: expand right edge of the full-size rows to 2 * cw_blocks * 8, bottom: replicate last row to even count, then h2v2 with bias 1,2,..
    def down(c):
        c2 = pad_edge(c, H + (H & 1), cw_blocks * 16)
        s = c2[0::2, 0::2] + c2[0::2, 1::2] + c2[1::2, 0::2] + c2[1::2, 1::2]
        bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]
        d = (s + bias) >> 2
        return pad_edge(d, ch_blocks * 8, cw_blocks * 8)
    CB, CR = down(cb), down(cr)
    def blocks_of(p, hb, wb, qt):
        b = p.reshape(hb, 8, wb, 8).transpose(0, 2, 1, 3) - 128
        c = fdct_islow(b).reshape(hb, wb, 64)
        a = np.abs(c); qv = qt[None, None, :] * 8
        quant = (a + (qv >> 1)) // qv
        return np.where(c < 0, -quant, quant)
    qY, qC = q[0], q[1]
    BY = blocks_of(Y, yh_blocks, yw_blocks, qY); BCB = blocks_of(CB, ch_blocks, cw_blocks, qC); BCR = blocks_of(CR, ch_blocks, cw_blocks, qC)
    dcY = huff_codes(*huff[(0, 0)]); acY = huff_codes(*huff[(1, 0)]); dcC = huff_codes(*huff[(0, 1)]); acC = huff_codes(*huff[(1, 1)])
    out = bytearray(); acc = 0; nb = 0
    def emit(code, size):
        nonlocal acc, nb
        acc = (acc << size) | (code & ((1 << size) - 1)); nb += size
        while nb >= 8:
            byte = (acc >> (nb - 8)) & 0xFF; out.append(byte)
            if byte == 0xFF: out.append(0)
            nb -= 8
        acc &= (1 << nb) - 1
    def nbits(v):
        return int(v).bit_length()
    def enc_block(blk, pred, dc, ac):
        diff = int(blk[0]) - pred
        t = diff; t2 = diff
        if t < 0: t = -t; t2 -= 1
        n = nbits(t)
        emit(dc[0][n], dc[1][n])
        if n: emit(t2, n)
        r = 0
        zz = blk[ZIGZAG]
        for k in range(1, 64):
            v = int(zz[k])
            if v == 0: r += 1; continue
            while r > 15:
                emit(ac[0][0xF0], ac[1][0xF0]); r -= 16
            t = v; t2 = v
            if t < 0: t = -t; t2 -= 1
            n = nbits(t)
            emit(ac[0][(r << 4) + n], ac[1][(r << 4) + n]); emit(t2, n); r = 0
        if r > 0: emit(ac[0][0], ac[1][0])
        return int(blk[0])
    pY = pCb = pCr = 0
    zero = np.zeros(64, np.int64)
    for my in range(mcuy):
        for mx in range(mcux):
            for by in range(2):
                for bx in range(2):
                    r_, c_ = 2 * my + by, 2 * mx + bx
                    if r_ < yh_blocks and c_ < yw_blocks: blk = BY[r_, c_]
                    else:
                        blk = zero.copy(); blk[0] = pY          # dummy block: DC of the previous block, zero AC
                    pY = enc_block(blk, pY, dcY, acY)
            for B, which in ((BCB, 0), (BCR, 1)):
                if my < ch_blocks and mx < cw_blocks: blk = B[my, mx]
                else:
                    blk = zero.copy(); blk[0] = pCb if which == 0 else pCr
                if which == 0: pCb = enc_block(blk, pCb, dcC, acC)
                else: pCr = enc_block(blk, pCr, dcC, acC)
    if nb: emit(0x7F, 7 if nb else 0)
    # flush: fill the last byte with ones
    return bytes(template[:sos_end]) + bytes(out) + b"\xff\xd9"

