"""Frame batcher (input side of the video loop): bit-exact against the reference's eager steps
(text_track_visualizer.py:313-321 + gom_lstmatcher.py:159-170 + ImageList.from_tensors padding)."""
import numpy as np
import pytest
import torch

MEAN = [123.675, 116.280, 103.530]      # configs/GoMatching_ICDAR15.yaml:4-5
STD = [58.395, 57.120, 57.375]


def eager_reference(frames_np, flip, div):
    """The reference's steps on the CPU, frame by frame (N = 1 per forward), stacked."""
    mean = torch.tensor(MEAN).view(3, 1, 1)
    std = torch.tensor(STD).view(3, 1, 1)
    outs = []
    for x in frames_np:
        if flip:
            x = x[:, :, ::-1]
        t = torch.as_tensor(x.astype("float32").transpose(2, 0, 1))
        t = (t - mean) / std
        h, w = t.shape[-2:]
        hp, wp = ((h + div - 1) // div * div, (w + div - 1) // div * div) if div > 1 else (h, w)
        outs.append(torch.nn.functional.pad(t, (0, wp - w, 0, hp - h), value=0.0))
    return torch.stack(outs)


def test_cpu_tensors_raise_and_shapes_are_checked():
    from gomatching_b200.video import batch_frames, padded_size
    assert padded_size(720, 1280, 32) == (736, 1280) and padded_size(1000, 1778, 32) == (1024, 1792)
    assert padded_size(5, 7, 0) == (5, 7)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        batch_frames(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), MEAN, STD)
    with pytest.raises(ValueError):
        batch_frames(torch.zeros(1, 4, 4, 3), MEAN, STD)
    with pytest.raises(ValueError):
        batch_frames(torch.zeros(1, 4, 4, 4, dtype=torch.uint8), MEAN, STD)


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,flip,div", [(2, 720, 1280, False, 32), (1, 720, 1280, True, 32), (3, 37, 53, True, 32),
                                            (1, 37, 53, False, 0), (2, 64, 96, True, 0), (1, 1000, 1778, True, 32),
                                            (1, 1, 1, False, 4)])
def test_batch_is_bit_identical_to_the_eager_steps(n, h, w, flip, div):
    from gomatching_b200.video import batch_frames
    rng = np.random.RandomState(h * 7 + w)
    frames = rng.randint(0, 256, size=(n, h, w, 3)).astype(np.uint8)
    frames[0, 0, 0] = (0, 128, 255)
    got = batch_frames(torch.from_numpy(frames).cuda(), MEAN, STD, flip_channels=flip, size_divisibility=div)
    ref = eager_reference(frames, flip, div)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert torch.equal(got.cpu(), ref)
    # the same eager ops on the device (what the reference actually runs) agree too
    dev = ((torch.from_numpy(frames[:, :, :, ::-1].copy() if flip else frames).cuda().permute(0, 3, 1, 2).float()
            - torch.tensor(MEAN, device="cuda").view(1, 3, 1, 1)) / torch.tensor(STD, device="cuda").view(1, 3, 1, 1))
    assert torch.equal(got[:, :, :h, :w], dev)


@pytest.mark.gpu
def test_single_frame_and_out_buffer():
    from gomatching_b200.video import batch_frames
    f = torch.randint(0, 256, (48, 80, 3), dtype=torch.uint8, device="cuda")
    a = batch_frames(f, MEAN, STD)
    out = torch.full((1, 3, 48, 80), 7.0, device="cuda")
    b = batch_frames(f, MEAN, STD, out=out)
    assert b.data_ptr() == out.data_ptr() and torch.equal(a, b)
    with pytest.raises(ValueError):
        batch_frames(f, MEAN, STD, out=torch.empty(1, 3, 48, 81, device="cuda"))
