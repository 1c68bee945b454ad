"""Row f4 on the GPU: image resize + normalise and RLE mask decode + nearest resize, bit-exact against the oracle (which is
pinned to Pillow / PyTorch in tests/test_input_oracle.py) and, at the real 1024 -> 1008 size, against Pillow itself."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as IO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,out", [(64, 80, 50), (37, 53, 100), (120, 90, 48), (33, 47, 94), (100, 100, 100)])
def test_image_resize_normalize_is_bit_exact(h, w, out):
    from sam3_lora_b200.data import GpuPreprocessor

    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = IO.to_tensor_normalize(IO.resize_bilinear_u8(img, out, out))
    prep = GpuPreprocessor(out)
    got = prep.image(img)
    assert got.shape == (3, out, out) and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), ref)
    assert np.array_equal(prep.image(torch.from_numpy(img).cuda()).cpu().numpy(), ref)      # device-resident input, cached tables


def test_sam3_size_image_matches_pillow_and_torchvision_arithmetic():
    from PIL import Image

    from sam3_lora_b200.data import GpuPreprocessor

    rng = np.random.default_rng(0)
    prep = GpuPreprocessor(1008)
    for (h, w) in ((1024, 1024), (480, 640), (1500, 2100)):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref8 = np.asarray(Image.fromarray(img).resize((1008, 1008), Image.BILINEAR)).copy()
        t = torch.from_numpy(ref8).permute(2, 0, 1).to(torch.float32).div(255)
        t = t.sub_(torch.full((3, 1, 1), 0.5)).div_(torch.full((3, 1, 1), 0.5))
        assert torch.equal(prep.image(img).cpu(), t), (h, w)


def test_rle_masks_are_bit_exact_and_batched():
    from sam3_lora_b200.data import GpuPreprocessor

    rng = np.random.default_rng(5)
    prep = GpuPreprocessor(100)
    rles, refs = [], []
    for (h, w) in ((40, 60), (100, 50), (50, 50), (120, 77), (1, 9), (200, 200)):
        runs, left = [], h * w
        while left > 0:
            c = int(min(left, rng.integers(0, 3 * h)))
            runs.append(c)
            left -= c
        if len(runs) % 3 == 0:
            runs = runs[:-1]                      # a run list may stop early: the rest of the image is background
        rles.append((runs, h, w))
        refs.append(IO.rle_mask_resized(runs, h, w, 100))
    got = prep.rle_masks(rles)
    assert got.dtype == torch.bool and got.shape == (len(rles), 100, 100)
    for g, r in zip(got.cpu().numpy(), refs):
        assert np.array_equal(g, r)
    assert prep.rle_masks([]).shape == (0, 100, 100)
    # SAM3 size against PyTorch's own nearest interpolation
    h, w = 768, 1024
    m = (rng.random((h, w)) > 0.5).astype(np.uint8)
    flat = m.T.reshape(-1)                                           # column-major
    change = np.flatnonzero(np.diff(flat)) + 1
    runs = np.diff(np.concatenate([[0], change, [flat.size]])).tolist()
    if flat[0] == 1:
        runs = [0] + runs
    ref = torch.nn.functional.interpolate(torch.from_numpy(m).float()[None, None], size=(1008, 1008), mode="nearest")[0, 0] > 0.5
    assert torch.equal(GpuPreprocessor(1008).rle_masks([(runs, h, w)])[0].cpu(), ref)
