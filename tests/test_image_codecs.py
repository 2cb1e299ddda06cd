"""The glTF front end's image decoders against stb_image itself (the reference decodes every texture with
stbi_load(..., STBI_rgb_alpha), MaterialUtils.h:226-249; the oracle links the reference's vendored stb_image and exposes it as
SailorPt_DecodeImage).  A texel is an input of the hot path, so the bar is byte equality.  Files are generated here with Pillow / OpenCV
(both in the image); host-only code, no GPU needed -- the product library still loads (and is the thing tested) without a CUDA device."""
import io
import os

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def product():
    import sailor_b200
    from sailor_b200 import build as product_build
    product_build.build()
    return sailor_b200.library()


def _picture(w, h, seed):
    """Smooth gradients + an edge + noise: exercises DC prediction, long AC runs, clamping and chroma upsampling."""
    r = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 7.0 + seed), 127 + 120 * np.cos(y / 5.0), 255.0 * ((x + y) % 32 < 16)], axis=-1)
    img += r.normal(0, 12, img.shape)
    img[h // 3:h // 2, w // 4:w // 2] = (250, 5, 5)
    return np.clip(img, 0, 255).astype(np.uint8)


def _jpeg(img, **kw):
    b = io.BytesIO()
    PIL.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


JPEG_CASES = [
    ("baseline 4:2:0", dict(quality=85, subsampling=2)),
    ("baseline 4:2:2", dict(quality=70, subsampling=1)),
    ("baseline 4:4:4", dict(quality=95, subsampling=0)),
    ("baseline optimised huffman", dict(quality=60, subsampling=2, optimize=True)),
    ("progressive 4:2:0", dict(quality=85, subsampling=2, progressive=True)),
    ("progressive 4:4:4", dict(quality=92, subsampling=0, progressive=True)),
    ("progressive 4:2:2 low quality", dict(quality=25, subsampling=1, progressive=True)),
    ("quality 100", dict(quality=100, subsampling=0)),
    ("quality 5", dict(quality=5, subsampling=2)),
]


@pytest.mark.parametrize("size", [(64, 48), (67, 35), (1, 1), (9, 17), (16, 8), (8, 16), (130, 3)])
@pytest.mark.parametrize("name,kw", JPEG_CASES)
def test_jpeg_decodes_byte_identical_to_stb_image(product, oracle, name, kw, size):
    data = _jpeg(_picture(size[0], size[1], len(name)), **kw)
    ref = oracle.decode_image(data)
    got = product.decode_image(data)
    assert got.shape == ref.shape == (size[1], size[0], 4)
    assert np.array_equal(got, ref), "%s %s: %d bytes differ, max |d| %d" % (name, size, int((got != ref).sum()), int(np.abs(got.astype(int) - ref.astype(int)).max()))


def test_jpeg_grey_cmyk_and_restart_intervals(product, oracle):
    img = _picture(75, 50, 3)
    files = {}
    b = io.BytesIO(); PIL.fromarray(img[..., 0]).save(b, "JPEG", quality=80); files["grey"] = b.getvalue()
    b = io.BytesIO(); PIL.fromarray(img[..., 0]).save(b, "JPEG", quality=80, progressive=True); files["grey progressive"] = b.getvalue()
    b = io.BytesIO(); PIL.fromarray(img).convert("CMYK").save(b, "JPEG", quality=80); files["cmyk (Adobe APP14)"] = b.getvalue()
    b = io.BytesIO(); PIL.fromarray(img).save(b, "JPEG", quality=80, restart_marker_blocks=3); files["restart every 3 MCUs"] = b.getvalue()
    b = io.BytesIO(); PIL.fromarray(img).save(b, "JPEG", quality=80, restart_marker_rows=1, progressive=True); files["progressive + restart per row"] = b.getvalue()
    b = io.BytesIO(); PIL.fromarray(img).save(b, "JPEG", quality=80, subsampling="4:1:1"); files["4:1:1 (nearest upsampling)"] = b.getvalue()
    try:
        import cv2
        ok, enc = cv2.imencode(".jpg", img[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 77, cv2.IMWRITE_JPEG_RST_INTERVAL, 5])
        if ok:
            files["opencv + DRI 5"] = enc.tobytes()
    except ImportError:
        pass
    for name, data in files.items():
        ref = oracle.decode_image(data)
        got = product.decode_image(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), "%s: %d bytes differ" % (name, int((got != ref).sum()))


def test_png_variants_decode_byte_identical_to_stb_image(product, oracle):
    """Every colour type and bit depth, interlaced (Adam7) and not, tRNS, widths smaller than an Adam7 pass."""
    img = _picture(37, 21, 9)
    rgba = np.concatenate([img, (255 - img[..., :1])], axis=-1)
    cases = {}
    for interlace in (False, True):
        tag = " interlaced" if interlace else ""
        for mode, arr in (("RGB", img), ("RGBA", rgba), ("L", img[..., 0]), ("LA", np.stack([img[..., 0], img[..., 1]], axis=-1)), ("P", img), ("1", img[..., 2] > 127)):
            im = PIL.fromarray(arr).convert(mode) if mode in ("P", "1") else PIL.fromarray(arr, mode)
            b = io.BytesIO(); im.save(b, "PNG", interlace=interlace) if False else None
            cases[mode + tag] = _png_bytes(im, interlace)
        im16 = PIL.fromarray((img[..., 0].astype(np.uint16) * 257 + 13).astype(np.uint16))
        cases["I;16" + tag] = _png_bytes(im16, interlace)
    for w, h in ((1, 1), (2, 3), (5, 1), (3, 9)):
        cases["tiny %dx%d interlaced" % (w, h)] = _png_bytes(PIL.fromarray(_picture(w, h, w + h)), True)
    try:
        import cv2
        big16 = (np.random.RandomState(1).randint(0, 65536, (19, 23, 3))).astype(np.uint16)
        ok, enc = cv2.imencode(".png", big16)
        if ok:
            cases["rgb 16-bit (opencv)"] = enc.tobytes()
    except ImportError:
        pass
    for name, data in cases.items():
        ref = oracle.decode_image(data)
        got = product.decode_image(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), "%s: %d bytes differ" % (name, int((got != ref).sum()))


def _png_bytes(im, interlace):
    """Pillow cannot write interlaced PNGs: Adam7 files are assembled here from the image's own passes."""
    if not interlace:
        b = io.BytesIO(); im.save(b, "PNG"); return b.getvalue()
    import struct, zlib
    b = io.BytesIO(); im.save(b, "PNG")
    src = b.getvalue()
    # parse the non-interlaced file Pillow wrote, re-pack its scanlines as Adam7
    off, chunks = 8, []
    while off < len(src):
        n = struct.unpack(">I", src[off:off + 4])[0]
        chunks.append((src[off + 4:off + 8], src[off + 8:off + 8 + n])); off += 12 + n
    ihdr = dict(chunks)[b"IHDR"]
    w, h, depth, color = struct.unpack(">IIBB", ihdr[:10])
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color]
    bpp_bits = ch * depth
    raw = zlib.decompress(b"".join(d for t, d in chunks if t == b"IDAT"))
    row_bytes = (w * bpp_bits + 7) // 8
    # undo the filters to get plain rows
    rows, prev = [], bytearray(row_bytes)
    bpp = max(1, bpp_bits // 8)
    p = 0
    for _ in range(h):
        f = raw[p]; cur = bytearray(raw[p + 1:p + 1 + row_bytes]); p += 1 + row_bytes
        for x in range(row_bytes):
            a = cur[x - bpp] if x >= bpp else 0; bb = prev[x]; c = prev[x - bpp] if x >= bpp else 0
            if f == 1: cur[x] = (cur[x] + a) & 255
            elif f == 2: cur[x] = (cur[x] + bb) & 255
            elif f == 3: cur[x] = (cur[x] + ((a + bb) >> 1)) & 255
            elif f == 4:
                pa, pb, pc = abs(bb - c), abs(a - c), abs(a + bb - 2 * c)
                cur[x] = (cur[x] + (a if pa <= pb and pa <= pc else (bb if pb <= pc else c))) & 255
        rows.append(bytes(cur)); prev = cur

    def get_px(row, x):          # bits of pixel x as an int
        if bpp_bits >= 8:
            return row[x * bpp_bits // 8:(x + 1) * bpp_bits // 8]
        bit = x * bpp_bits
        return (row[bit >> 3] >> (8 - bpp_bits - (bit & 7))) & ((1 << bpp_bits) - 1)
    out = bytearray()
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        xs, ys = range(x0, w, dx), range(y0, h, dy)
        if not len(xs) or not len(ys):
            continue
        for y in ys:
            out.append(0)
            if bpp_bits >= 8:
                out += b"".join(get_px(rows[y], x) for x in xs)
            else:
                acc, nb = 0, 0
                line = bytearray()
                for x in xs:
                    acc = (acc << bpp_bits) | get_px(rows[y], x); nb += bpp_bits
                    if nb == 8:
                        line.append(acc); acc, nb = 0, 0
                if nb:
                    line.append(acc << (8 - nb))
                out += line

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    res = src[:8] + chunk(b"IHDR", ihdr[:12] + b"\x01")
    for t, d in chunks:
        if t in (b"PLTE", b"tRNS"):
            res += chunk(t, d)
    return res + chunk(b"IDAT", zlib.compress(bytes(out))) + chunk(b"IEND", b"")


def test_a_gltf_with_jpeg_textures_loads_and_matches_the_oracle(product, oracle, tmp_path):
    """a JPEG-textured file used to fail with ERR_FORMAT.  The same scene with its base-colour texture as a progressive JPEG imports to
    the same texels as the oracle's stb_image path (host-side check through SailorPt_DecodeImage; the scene load itself needs a device)."""
    import scenes
    data = _jpeg(_picture(64, 64, 4), quality=90, progressive=True)
    assert np.array_equal(product.decode_image(data), oracle.decode_image(data))
    bad = b"\x00\x01\x02not an image at all"
    with pytest.raises(Exception):
        product.decode_image(bad)


def test_damaged_files_fail_or_decode_like_stb_image(product, oracle):
    """Truncated files and files with damaged entropy-coded / compressed data: the decoders reject what stb_image v2.27 (the version the reference's
    path tracer includes) rejects and otherwise produce the same pixels.  Exceptions are allowed for at most 2 % of the damaged JPEGs: stb's SSE2
    inverse DCT wraps around in 16 bits on coefficients no valid stream contains, where this decoder's integer IDCT (= stb's scalar one) does not."""
    r = np.random.RandomState(1)
    img = _picture(40, 28, 1)
    seeds = {"baseline": _jpeg(img, quality=80, subsampling=2), "progressive": _jpeg(img, quality=80, subsampling=0, progressive=True),
             "restart": _jpeg(img, quality=60, subsampling=1, restart_marker_blocks=2), "png": _png_bytes(PIL.fromarray(img), False), "png adam7": _png_bytes(PIL.fromarray(img), True)}

    def data_start(d):
        if d[:2] != b"\xff\xd8":
            return d.index(b"IDAT") + 4
        i = 2
        while True:
            m, n = d[i + 1], (d[i + 2] << 8) | d[i + 3]
            if m == 0xDA:
                return i + 2 + n
            i += 2 + n

    def outcome(lib, d):
        try:
            return lib.decode_image(d)
        except Exception:
            return None
    mismatches, total, truncated_mismatches = 0, 0, 0
    for name, seed in seeds.items():
        s0 = data_start(seed)
        for it in range(400):
            d = bytearray(seed)
            truncated = it % 4 == 0
            if truncated:
                d = d[:r.randint(s0, len(d))]
            else:
                for _ in range(r.randint(1, 4)):
                    d[r.randint(s0, len(d) - (12 if name.startswith("png") else 0))] = r.randint(0, 256)
            a, b = outcome(product, bytes(d)), outcome(oracle, bytes(d))
            same = (a is None and b is None) or (a is not None and b is not None and a.shape == b.shape and np.array_equal(a, b))
            total += 1
            mismatches += 0 if same else 1
            truncated_mismatches += 0 if (same or not truncated) else 1
    assert truncated_mismatches == 0, "a truncated file decoded differently"
    assert mismatches <= 0.02 * total, (mismatches, total)
