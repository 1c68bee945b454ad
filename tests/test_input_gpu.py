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


def _random_polygon(rng, h, w, k):
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    r = rng.uniform(0.15, 0.6, k) * min(h, w)
    cx, cy = rng.uniform(0.2, 0.8) * w, rng.uniform(0.2, 0.8) * h
    pts = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1)          # may leave the image: clipping is exercised
    return np.round(pts, 2).reshape(-1).tolist()


def test_polygon_masks_are_bit_exact_with_the_rleFrPoly_restatement():
    """frPyObjects + merge + decode + nearest resize on the GPU (crossing kernel -> sort -> parity fill) against the oracle:
    star-shaped and self-intersecting polygons, several polygons per object, polygons that leave the image, short lists."""
    from sam3_lora_b200.data import GpuPreprocessor

    rng = np.random.default_rng(11)
    prep = GpuPreprocessor(96)
    objs = []
    for (h, w) in ((40, 50), (64, 64), (90, 33), (120, 200)):
        for n_poly in (1, 2):
            polys = [_random_polygon(rng, h, w, int(rng.integers(3, 12))) for _ in range(n_poly)]
            objs.append((polys, h, w))
    objs.append(([[10, 10, 20, 10, 20, 20, 10, 20]], 40, 50))
    objs.append(([[1, 1, 2, 2], [5, 5, 30, 8, 12, 33]], 40, 50))                  # first list is degenerate and skipped
    objs.append(([[3, 3, 30, 30, 30, 3, 3, 30]], 40, 40))                         # bow-tie (self-intersecting)
    objs.append(([], 20, 20))                                                     # object without polygons
    got = prep.polygon_masks(objs).cpu().numpy()
    assert got.shape == (len(objs), 96, 96) and got.dtype == np.bool_
    for i, (polys, h, w) in enumerate(objs):
        ref = IO.poly_mask_resized(polys, h, w, 96)
        assert np.array_equal(got[i], ref), (i, h, w, int((got[i] ^ ref).sum()))
    assert prep.polygon_masks([]).shape == (0, 96, 96)
    # SAM3 size: one 1024 x 1024 annotation with a 40-vertex outline
    h = w = 1024
    poly = _random_polygon(rng, h, w, 40)
    big = GpuPreprocessor(1008).polygon_masks([([poly], h, w)])[0].cpu().numpy()
    assert np.array_equal(big, IO.poly_mask_resized([poly], h, w, 1008))


def test_nvjpeg_decode_feeds_the_resize_kernel_and_is_close_to_pil():
    import io

    from PIL import Image

    from sam3_lora_b200.data import GpuPreprocessor

    yy, xx = np.mgrid[0:240, 0:320]
    img = np.stack([(xx * 255 // 319), (yy * 255 // 239), ((xx + yy) % 256)], -1).astype(np.uint8)     # smooth: JPEG-friendly
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=95, subsampling=0)
    prep = GpuPreprocessor(128)
    dec = prep.decode_jpeg(buf.getvalue())
    assert dec.is_cuda and dec.dtype == torch.uint8 and dec.shape == (240, 320, 3)
    pil = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    diff = np.abs(dec.cpu().numpy().astype(int) - pil.astype(int))
    assert diff.max() <= 4 and diff.mean() < 0.6           # different IDCT implementations: a few grey levels, not bit-exact
    x = prep.image(dec)
    ref = torch.from_numpy(IO.to_tensor_normalize(IO.resize_bilinear_u8(dec.cpu().numpy(), 128, 128)))
    assert torch.equal(x.cpu(), ref)                        # from the decoded pixels on, the pipeline is exact


def test_prefetcher_overlaps_and_preserves_order():
    from sam3_lora_b200.data import Prefetcher

    def make(i):
        return {"i": i, "x": torch.full((1024,), float(i)).pin_memory()}

    seen = []
    for item in Prefetcher((make(i) for i in range(6)), depth=2, device="cuda",
                           transform=lambda d: {"i": d["i"], "x": d["x"].to("cuda", non_blocking=True)}):
        assert item["x"].is_cuda and float(item["x"].sum()) == 1024.0 * item["i"]
        seen.append(item["i"])
    assert seen == list(range(6))
    with pytest.raises(ZeroDivisionError):
        list(Prefetcher((1 // (3 - i) for i in range(6)), device="cuda"))
