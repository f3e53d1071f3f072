"""Image ingest: the library's PNG decoder / KITTI-layout reader against what the reference's
loadImageLeft/Right compute (cv::imread(IMREAD_COLOR) + cvtColor(BGR2GRAY), reference src/utils.cpp:172-190) -- here
cv2 4.13 on the same files.  Host-only (the C-ABI library loads without a GPU)."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cv2 = pytest.importorskip("cv2")
from visual_odom_b200 import capi  # noqa: E402


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def _encode(rows, w, h, depth, ctype, bpp, filters, plte=None, split=1):
    """A from-scratch PNG writer so every filter type / colour type / bit depth is exercised (rows: bytes per scanline)."""
    out = bytearray()
    prev = bytes(len(rows[0]))
    for y, row in enumerate(rows):
        ft = filters[y % len(filters)]
        f = bytearray(len(row))
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            f[i] = (v - pred) & 255
        out.append(ft); out += f
        prev = row
    z = zlib.compress(bytes(out), 6)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if plte is not None:
        png += _chunk(b"PLTE", plte)
    k = max(1, len(z) // split)
    for i in range(0, len(z), k):
        png += _chunk(b"IDAT", z[i:i + k])
    return png + _chunk(b"IEND", b"")


def _cv_ref(data):
    bgr = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    return bgr, cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)


def test_gray_formula_is_cv2s_over_the_whole_cube(tmp_path):
    a = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(4096, 4096, 3)     # every (b, g, r)
    ok, enc = cv2.imencode(".png", cube[:, :, ::-1].copy()[:, :, ::-1], [cv2.IMWRITE_PNG_COMPRESSION, 1])
    assert ok
    bgr, gray = capi.png_decode(enc.tobytes())
    assert np.array_equal(bgr, cube)
    assert np.array_equal(gray, cv2.cvtColor(cube, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("kind", ["gray8", "gray16", "bgr8", "bgr16", "bgra8"])
def test_decode_cv2_written_files(kind):
    rng = np.random.default_rng(3)
    h, w = 137, 251
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 4)).astype(np.float32), (0, 0), 2.0)
    if kind == "gray8": img = smooth[..., 0].astype(np.uint8)
    elif kind == "gray16": img = (smooth[..., 0] * 257).astype(np.uint16)
    elif kind == "bgr8": img = smooth[..., :3].astype(np.uint8)
    elif kind == "bgr16": img = (smooth[..., :3] * 257).astype(np.uint16)
    else: img = smooth.astype(np.uint8)
    ok, enc = cv2.imencode(".png", img)
    assert ok
    data = enc.tobytes()
    bgr, gray = capi.png_decode(data)
    rb, rg = _cv_ref(data)
    assert np.array_equal(bgr, rb) and np.array_equal(gray, rg)
    if kind == "gray8":
        assert np.array_equal(gray, img)          # imread(COLOR) + cvtColor is the identity on gray files (KITTI)


@pytest.mark.parametrize("ctype,depth", [(0, 1), (0, 2), (0, 4), (0, 8), (0, 16), (2, 8), (2, 16), (3, 1), (3, 2), (3, 4), (3, 8),
                                         (4, 8), (4, 16), (6, 8), (6, 16)])
def test_every_colour_type_bit_depth_and_filter(ctype, depth):
    rng = np.random.default_rng(100 * ctype + depth)
    w, h = 53, 41
    ns = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bits = ns * depth
    rowbytes = (w * bits + 7) // 8
    rows = [bytes(rng.integers(0, 256, rowbytes, dtype=np.uint8)) for _ in range(h)]
    plte = bytes(rng.integers(0, 256, 3 * (1 << depth), dtype=np.uint8)) if ctype == 3 else None
    data = _encode(rows, w, h, depth, ctype, max(1, bits // 8), [0, 1, 2, 3, 4, 4, 3, 1], plte, split=3)
    assert capi.png_info(data) == (w, h, ctype, depth)
    bgr, gray = capi.png_decode(data)
    rb, rg = _cv_ref(data)
    assert np.array_equal(bgr, rb), "BGR differs from imread(IMREAD_COLOR)"
    assert np.array_equal(gray, rg), "gray differs from cvtColor(BGR2GRAY)"


def test_bad_files_are_rejected():
    rows = [bytes(16) for _ in range(4)]
    good = _encode(rows, 16, 4, 8, 0, 1, [0])
    with pytest.raises(RuntimeError, match="not a PNG"):
        capi.png_decode(b"JFIF" + good[4:])
    bad_crc = bytearray(good); bad_crc[-20] ^= 1
    with pytest.raises(RuntimeError, match="CRC|corrupt|truncated"):
        capi.png_decode(bytes(bad_crc))
    with pytest.raises(RuntimeError, match="truncated|corrupt"):
        capi.png_decode(good[:len(good) // 2])
    inter = bytearray(good); inter[28] = 1          # IHDR interlace byte
    inter[29:33] = struct.pack(">I", zlib.crc32(bytes(inter[12:29])) & 0xffffffff)
    with pytest.raises(RuntimeError, match="interlac"):
        capi.png_decode(bytes(inter))


def _write_sequence(root, frames, first=0):
    os.makedirs(os.path.join(root, "image_0")); os.makedirs(os.path.join(root, "image_1"))
    for i, (l, r) in enumerate(frames):
        assert cv2.imwrite(os.path.join(root, "image_0", "%06d.png" % (first + i)), l)
        assert cv2.imwrite(os.path.join(root, "image_1", "%06d.png" % (first + i)), r)


@pytest.mark.parametrize("colour", [False, True])
def test_reader_kitti_layout_prefetch(tmp_path, colour):
    rng = np.random.default_rng(9)
    shape = (94, 311, 3) if colour else (94, 311)
    frames = [(rng.integers(0, 256, shape, dtype=np.uint8), rng.integers(0, 256, shape, dtype=np.uint8)) for _ in range(11)]
    _write_sequence(str(tmp_path), frames, first=5)
    rd = capi.SequenceReader(str(tmp_path), 5, 11, threads=3, depth=3)
    for i in range(11):
        l, r, fid = rd.next()
        assert fid == 5 + i
        assert np.array_equal(l, frames[i][0]) and np.array_equal(r, frames[i][1])
    with pytest.raises(RuntimeError, match="end of sequence"):
        rd.next()
    rd.close()
    # forced gray delivery of colour files == cvtColor on the host
    rd = capi.SequenceReader(str(tmp_path), 5, 4, threads=2, depth=2, force_channels=1)
    for i in range(4):
        l, r, fid = rd.next()
        ref = frames[i][0] if not colour else cv2.cvtColor(frames[i][0], cv2.COLOR_BGR2GRAY)
        assert np.array_equal(l, ref)
    rd.close()


def test_reader_reports_missing_files(tmp_path):
    rng = np.random.default_rng(1)
    frames = [(rng.integers(0, 256, (32, 48), dtype=np.uint8),) * 2 for _ in range(3)]
    _write_sequence(str(tmp_path), frames)
    with pytest.raises(RuntimeError, match="cannot open"):
        capi.SequenceReader(str(tmp_path / "nope"), 0, 3)
    rd = capi.SequenceReader(str(tmp_path), 0, 5, threads=2, depth=2)      # frames 3, 4 do not exist
    for _ in range(3):
        rd.next()
    with pytest.raises(RuntimeError, match="cannot open"):
        rd.next()
    rd.close()


def test_facade_load_image_left_right(tmp_path):
    """compat/utils.h loadImageLeft / loadImageRight (reference signatures) against cv2.imread + cvtColor."""
    import subprocess
    from visual_odom_b200 import build
    build.build_native(); build.build_facade()
    rng = np.random.default_rng(4)
    frames = [(rng.integers(0, 256, (60, 85, 3), dtype=np.uint8), rng.integers(0, 256, (60, 85), dtype=np.uint8))]
    _write_sequence(str(tmp_path), frames, first=7)
    exe = os.path.join(ROOT, "tests", "cpp", "utils_main")
    out = tmp_path / "o.bin"
    r = subprocess.run([exe, "load", str(tmp_path) + "/", "7", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    off = 0
    for cam in range(2):
        w, h = struct.unpack_from("<ii", raw, off); off += 8
        col = np.frombuffer(raw, np.uint8, 3 * w * h, off).reshape(h, w, 3); off += 3 * w * h
        gray = np.frombuffer(raw, np.uint8, w * h, off).reshape(h, w); off += w * h
        ref = cv2.imread(os.path.join(str(tmp_path), "image_%d" % cam, "000007.png"), cv2.IMREAD_COLOR)
        assert np.array_equal(col, ref) and np.array_equal(gray, cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY))
    r = subprocess.run([exe, "load", str(tmp_path) + "/", "8", str(out)], capture_output=True, text=True)
    assert r.returncode == 6 and "cannot open" in r.stderr


@pytest.mark.parametrize("threads,depth", [(1, 3), (8, 3), (8, 16), (3, 5)])
def test_reader_ring_under_pressure(tmp_path, threads, depth):
    """More decoders than ring slots, a consumer that holds two frames (the contract vo_seq_submit relies on): frames come
    back in order, the two most recent hand-outs stay intact while later frames are being decoded, early close joins."""
    import ctypes as C
    rng = np.random.default_rng(threads * 100 + depth)
    n = 40
    frames = [(rng.integers(0, 256, (48, 64), dtype=np.uint8), rng.integers(0, 256, (48, 64), dtype=np.uint8)) for _ in range(n)]
    _write_sequence(str(tmp_path), frames)
    rd = capi.SequenceReader(str(tmp_path), 0, n, threads=threads, depth=depth)
    held = []
    for i in range(n):
        l, r, w, h, pitch, ch, fid = rd.next_ptr()
        assert (w, h, ch, fid) == (64, 48, 1, i)
        held.append((i, l, r))
        held = held[-2:]                                   # the consumer keeps using the last two hand-outs
        for j, lp, rp in held:
            la = np.ctypeslib.as_array((C.c_uint8 * (48 * 64)).from_address(lp)).reshape(48, 64)
            ra = np.ctypeslib.as_array((C.c_uint8 * (48 * 64)).from_address(rp)).reshape(48, 64)
            assert np.array_equal(la, frames[j][0]) and np.array_equal(ra, frames[j][1]), (i, j)
    rd.close()
    rd = capi.SequenceReader(str(tmp_path), 0, n, threads=threads, depth=depth)      # close with work outstanding
    rd.next_ptr()
    rd.close()
