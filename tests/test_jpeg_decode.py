"""JPEG frame decode (the data format before the path): the reference reads frames with
``read_image(path, format="BGR")`` (eval.py:324-327) = Pillow's libjpeg-turbo with default settings.

* not gpu: the CPU restatement ``oracle/jpeg_oracle.cpp`` (IJG jidctint / jdsample / jdcolor arithmetic + the product's
  host entropy decoder) is PINNED on Pillow itself -- every pixel equal to ``PIL.Image.open(...).convert("RGB")``.
* gpu: the device decoder (csrc/jpeg_decode.cu through the C ABI) equals Pillow (and hence the oracle) on the same files,
  through ``decode_jpeg`` / ``read_image_bgr`` and through ``ClipTracker``'s frame input.
"""
import io
import os

import numpy as np
import pytest
from PIL import Image


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
    img = np.array(Image.fromarray(a).resize((w, h), Image.BICUBIC)).astype(np.float32) + rng.normal(0, 10, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def _jpeg(img, mode="RGB", **kw):
    b = io.BytesIO()
    Image.fromarray(img).convert(mode).save(b, "JPEG", **kw)
    return b.getvalue()


def _pillow_rgb(data):
    return np.array(Image.open(io.BytesIO(data)).convert("RGB"))


SIZES = [(16, 16), (61, 47), (33, 70), (97, 131), (8, 8), (17, 23), (5, 3), (1, 1), (2, 4), (240, 320)]
CASES = [dict(quality=q, subsampling=s) for s in (0, 1, 2) for q in (2, 30, 75, 95, 100)] + [
    dict(quality=85, optimize=True), dict(quality=85, restart_marker_blocks=7),
    dict(quality=85, restart_marker_rows=1, subsampling=2), dict(quality=60, subsampling=1, restart_marker_blocks=3, optimize=True)]


def _files():
    out = []
    for k, (h, w) in enumerate(SIZES):
        img = _image(h, w, k)
        for kw in (CASES if h * w <= 97 * 131 else CASES[::4]):
            out.append(((h, w), kw, _jpeg(img, **kw)))
        out.append(((h, w), {"mode": "L"}, _jpeg(img, mode="L", quality=80)))
    sat = np.zeros((64, 64, 3), np.uint8)                    # saturated content: the range-limit tables
    sat[::2] = 255
    sat[:, ::3, 1] = 255
    out.append(((64, 64), {"saturated": 100}, _jpeg(sat, quality=100, subsampling=0)))
    out.append(((64, 64), {"saturated": 10}, _jpeg(sat, quality=10, subsampling=2)))
    return out


def test_oracle_is_pinned_on_pillow():
    from oracle import jpeg_oracle as J
    files = _files()
    assert len(files) > 100
    for size, kw, data in files:
        want = _pillow_rgb(data)
        got = J.decode_rgb(data)
        assert got.shape == want.shape and np.array_equal(got, want), (size, kw)


def test_oracle_full_frame_720p_and_error_codes():
    from oracle import jpeg_oracle as J
    data = _jpeg(_image(720, 1280, 99), quality=90, subsampling=2)
    assert np.array_equal(J.decode_rgb(data), _pillow_rgb(data))
    with pytest.raises(ValueError) as e:
        J.decode_rgb(_jpeg(_image(40, 40, 1), progressive=True))
    assert e.value.args[0] == 3                              # unsupported, like the product: no silent fallback
    with pytest.raises(ValueError) as e:
        J.decode_rgb(data[:len(data) // 2])
    assert e.value.args[0] == 1                              # truncated
    with pytest.raises(ValueError):
        J.decode_rgb(b"not a jpeg at all")


def test_host_side_of_the_product_decoder():
    """No GPU needed: frame size, EXIF orientation and the unsupported / CPU-tensor errors."""
    import torch
    from gomatching_b200.video import jpeg
    im = Image.fromarray(_image(33, 47, 5))
    ex = im.getexif()
    ex[0x0112] = 6
    b = io.BytesIO()
    im.save(b, "JPEG", exif=ex)
    assert jpeg.jpeg_size(b.getvalue()) == (33, 47, 3)
    assert jpeg.exif_orientation(b.getvalue()) == 6
    assert jpeg.exif_orientation(_jpeg(_image(8, 8, 0))) == 1
    assert jpeg.jpeg_size(_jpeg(_image(9, 11, 0), mode="L")) == (9, 11, 1)
    from gomatching_b200._native import MSDAError
    with pytest.raises(MSDAError):
        jpeg.jpeg_size(_jpeg(_image(40, 40, 1), progressive=True))
    with pytest.raises(RuntimeError):
        jpeg.decode_jpeg(_jpeg(_image(8, 8, 0)), torch.device("cpu"))      # no CPU path


@pytest.mark.gpu
def test_device_decoder_equals_pillow():
    import torch
    from gomatching_b200.video import jpeg
    from gomatching_b200._native import MSDAError
    for size, kw, data in _files():
        want = _pillow_rgb(data)
        got = jpeg.decode_jpeg(data, "cuda", bgr=False)
        assert got.dtype == torch.uint8 and tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want), (size, kw)
    data = _jpeg(_image(720, 1280, 99), quality=90, subsampling=2)
    bgr = jpeg.decode_jpeg(bytearray(data), "cuda")                        # default: BGR like read_image(format="BGR")
    assert np.array_equal(bgr.cpu().numpy(), _pillow_rgb(data)[:, :, ::-1])
    data1080 = _jpeg(_image(1080, 1920, 7), quality=75, subsampling=1)
    assert np.array_equal(jpeg.decode_jpeg(memoryview(data1080), "cuda", bgr=False).cpu().numpy(), _pillow_rgb(data1080))
    with pytest.raises(MSDAError):
        jpeg.decode_jpeg(data[:len(data) // 2], "cuda")                     # truncated
    with pytest.raises(MSDAError):
        jpeg.decode_jpeg(_jpeg(_image(40, 40, 1), progressive=True), "cuda")   # unsupported: raises, no fallback
    with pytest.raises(ValueError):
        jpeg.decode_jpeg(data, "cuda", out=torch.empty((10, 10, 3), dtype=torch.uint8, device="cuda"))


@pytest.mark.gpu
def test_read_image_bgr_matches_the_reference_frame_read(tmp_path):
    """detectron2's read_image: PIL open -> EXIF transpose -> RGB -> BGR (eval.py:327)."""
    from PIL import ImageOps
    from gomatching_b200.video import jpeg
    img = _image(45, 83, 3)
    for orientation in (None, 1, 2, 3, 4, 5, 6, 7, 8):
        im = Image.fromarray(img)
        path = os.path.join(tmp_path, "f%s.jpg" % orientation)
        if orientation is None:
            im.save(path, "JPEG", quality=88)
        else:
            ex = im.getexif()
            ex[0x0112] = orientation
            im.save(path, "JPEG", quality=88, exif=ex)
        want = np.asarray(ImageOps.exif_transpose(Image.open(path)).convert("RGB"))[:, :, ::-1]
        got = jpeg.read_image_bgr(path, "cuda").cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), orientation


def test_entropy_decoder_survives_corrupted_streams():
    """The host Huffman stage is shared by the oracle and the product: byte flips, stray 0xFF, deletions and insertions
    must end in a decoded image or an error code, never in a crash or an out-of-bounds block index."""
    from oracle import jpeg_oracle as J
    rng = np.random.default_rng(5)
    img = _image(67, 93, 11)
    base = [_jpeg(img, quality=80), _jpeg(img, quality=50, subsampling=0, optimize=True),
            _jpeg(img, quality=90, restart_marker_blocks=4), _jpeg(img, quality=70, subsampling=1)]
    outcomes = {"ok": 0}
    for it in range(2000):
        d = bytearray(base[it % len(base)])
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(2, len(d)))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                d[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                d[pos] = 0xFF
            elif mode == 2:
                del d[pos:pos + int(rng.integers(1, 20))]
            else:
                d[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 8))).astype(np.uint8))
        try:
            out = J.decode_rgb(bytes(d))
            assert out.ndim == 3 and out.shape[2] == 3
            outcomes["ok"] += 1
        except ValueError as e:
            assert e.args[0] in (1, 2, 3)
            outcomes[e.args[0]] = outcomes.get(e.args[0], 0) + 1
    assert outcomes["ok"] > 0 and outcomes.get(2, 0) > 0 and outcomes.get(1, 0) > 0
