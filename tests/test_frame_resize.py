"""Frame resize (the data format before the path): gomatching_b200/video/resize.py +
csrc/frame_resize.cu against Pillow itself -- the reference resizes test frames with
ResizeShortestEdge -> PIL.Image.resize(BILINEAR) (gomatching/text_track_visualizer.py:283-284, :318-319)."""
import numpy as np
import pytest
import torch
from PIL import Image

from gomatching_b200.video import resize as R

SIZES = [((72, 128), (100, 178)), ((72, 128), (50, 89)), ((37, 53), (37, 90)), ((64, 48), (23, 48)), ((45, 80), (144, 256)),
         ((100, 100), (33, 35)), ((9, 7), (40, 3))]


def pil_resize(a, new_h, new_w):
    if a.shape[2] == 1:
        return np.asarray(Image.fromarray(a[:, :, 0]).resize((new_w, new_h), Image.BILINEAR))[:, :, None]
    return np.asarray(Image.fromarray(a).resize((new_w, new_h), Image.BILINEAR))


def test_shortest_edge_rule():
    assert R.shortest_edge_size(720, 1280, 1000, 3000) == (1000, 1778)          # ICDAR15 720p -> SURVEY s8 table
    assert R.shortest_edge_size(1080, 1920, 1280, 3000) == (1280, 2276)         # DSText
    assert R.shortest_edge_size(1280, 720, 1000, 1500) == (1500, 844)           # long side capped
    assert R.shortest_edge_size(500, 500, 500, 3000) == (500, 500)


@pytest.mark.parametrize("src,dst", SIZES)
def test_coefficient_tables_reproduce_pillow(src, dst):
    """numpy emulation of the two passes with the restated tables == Pillow, bit for bit (no GPU needed)."""
    a = np.random.default_rng(src[0] * 131 + dst[1]).integers(0, 256, src + (3,), dtype=np.uint8)
    assert np.array_equal(R.resample_reference(a, *dst), pil_resize(a, *dst))
    edge = np.zeros(src + (3,), np.uint8)
    edge[::2] = 255                                                             # worst case for rounding: full-swing rows
    assert np.array_equal(R.resample_reference(edge, *dst), pil_resize(edge, *dst))


def test_table_shapes():
    b, k, ks = R.bilinear_coeffs(128, 178)
    assert ks == 3 and k.shape == (178, 3) and b.shape == (178, 2)
    assert (k.sum(1) - (1 << 22)).__abs__().max() <= 2                          # normalised 22-bit weights
    b, k, ks = R.bilinear_coeffs(1280, 500)                                     # shrinking: antialiased support 2.56
    assert ks == 7 and b[:, 1].max() <= 7


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", SIZES + [((720, 1280), (1000, 1778)), ((1080, 1920), (720, 1280))])
def test_gpu_resize_is_bit_identical_to_pillow(src, dst):
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, (2,) + src + (3,), dtype=np.uint8)
    out = R.resize_frames_u8(torch.from_numpy(frames).cuda(), *dst).cpu().numpy()
    assert out.shape == (2,) + dst + (3,)
    for i in range(2):
        assert np.array_equal(out[i], pil_resize(frames[i], *dst))
    one = R.resize_frames_u8(torch.from_numpy(frames[0]).cuda(), *dst)
    assert one.shape == dst + (3,) and np.array_equal(one.cpu().numpy(), out[0])


def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        R.resize_frames_u8(torch.zeros(4, 4, 3, dtype=torch.uint8), 8, 8)
