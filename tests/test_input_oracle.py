"""Row f4 on the CPU: the input-pipeline oracle is pinned bit-exactly against the Pillow and PyTorch installed in the build
container (the third-party code the reference calls), and the HOST coefficient routine of libsam3b against the oracle."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as IO

SIZES = [(64, 80, 50), (37, 53, 100), (120, 90, 48), (100, 100, 64), (33, 47, 94), (256, 200, 252)]


@pytest.mark.parametrize("h,w,out", SIZES)
def test_resize_oracle_is_bit_exact_with_pillow_and_torch_normalize(h, w, out):
    from PIL import Image

    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((out, out), Image.BILINEAR))
    got = IO.resize_bilinear_u8(img, out, out)
    assert np.array_equal(got, ref)
    # ToTensor + Normalize(0.5, 0.5) as torchvision does them (float32 ops)
    t = torch.from_numpy(ref.copy()).permute(2, 0, 1).to(torch.float32).div(255)
    t = t.sub_(torch.tensor([0.5, 0.5, 0.5]).view(3, 1, 1)).div_(torch.tensor([0.5, 0.5, 0.5]).view(3, 1, 1))
    assert np.array_equal(IO.to_tensor_normalize(got), t.numpy())


@pytest.mark.parametrize("in_size,out_size", [(1024, 1008), (640, 1008), (512, 1008), (1008, 1008), (2000, 1008), (53, 100), (7, 1008)])
def test_library_coefficients_equal_the_oracle(in_size, out_size):
    from sam3_lora_b200.data import resample_coeffs

    b, k, ks = resample_coeffs(in_size, out_size)
    rb, rk, rks = IO.resample_coeffs(in_size, out_size)
    assert ks == rks and np.array_equal(b, rb) and np.array_equal(k, rk)
    assert (k.sum(1) - (1 << IO.PRECISION_BITS)).__abs__().max() <= k.shape[1]     # rows are normalised up to rounding


def _random_rle(rng, h, w):
    runs, left = [], h * w
    while left > 0:
        c = int(min(left, rng.integers(0, 3 * h)))
        runs.append(c)
        left -= c
    return runs


@pytest.mark.parametrize("h,w,out", [(40, 60, 100), (100, 50, 100), (50, 50, 100), (120, 77, 84), (1, 9, 16)])
def test_rle_mask_oracle_matches_numpy_decode_and_torch_nearest(h, w, out):
    rng = np.random.default_rng(h + w)
    counts = _random_rle(rng, h, w)
    m = IO.rle_decode(counts, h, w)
    flat = np.concatenate([np.full(c, i & 1, np.uint8) for i, c in enumerate(counts)])[: h * w]
    assert np.array_equal(m, np.pad(flat, (0, h * w - len(flat))).reshape(w, h).T)
    ref = torch.nn.functional.interpolate(torch.from_numpy(m).float()[None, None], size=(out, out), mode="nearest")[0, 0] > 0.5
    assert np.array_equal(IO.rle_mask_resized(counts, h, w, out), ref.numpy())


def test_compressed_rle_string_round_trip():
    """rleFrString restated twice (oracle and sam3_lora_b200.data) against an encoder written from maskApi.c's rleToString."""
    from sam3_lora_b200.data import rle_counts

    def to_string(cnts):
        s = []
        for i, c in enumerate(cnts):
            x = int(c)
            if i > 2:
                x -= int(cnts[i - 2])
            more = True
            while more:
                ch = x & 0x1F
                x >>= 5
                more = (x != -1) if (ch & 0x10) else (x != 0)
                if more:
                    ch |= 0x20
                s.append(chr(ch + 48))
        return "".join(s)

    rng = np.random.default_rng(7)
    for _ in range(20):
        cnts = [int(v) for v in rng.integers(0, 5000, size=rng.integers(1, 40))]
        enc = to_string(cnts)
        assert IO.rle_from_string(enc) == cnts
        assert rle_counts(enc) == cnts and rle_counts(enc.encode()) == cnts
    assert rle_counts([3, 4, 5]) == [3, 4, 5]


def test_polygon_oracle_known_answers():
    """pycocotools is not installed, so rleFrPoly's restatement is checked against facts that hold for pycocotools itself:
    an axis-aligned integer box polygon [x, y, x+w, y+h] covers exactly the w*h pixels [x, x+w) x [y, y+h) (the
    frPyObjects / area identity every COCO tool relies on), vertex order and starting vertex do not matter, polygons are
    clipped to the image, several polygons of one object are OR-ed (mask_utils.merge), degenerate lists are ignored."""
    box = [10, 10, 20, 10, 20, 20, 10, 20]
    m = IO.poly_mask([box], 40, 50)
    assert m.shape == (40, 50) and m.sum() == 100 and m[10:20, 10:20].all()
    assert np.array_equal(IO.poly_mask([box[2:] + box[:2]], 40, 50), m)                   # rotated start vertex
    rev = [c for pt in reversed(list(zip(box[0::2], box[1::2]))) for c in pt]
    assert np.array_equal(IO.poly_mask([rev], 40, 50), m)                                 # clockwise / counter-clockwise
    big = IO.poly_mask([[5, 5, 60, 5, 60, 45, 5, 45]], 40, 50)
    assert big.sum() == 45 * 35 and big[5:, 5:].all() and not big[:5].any()
    two = IO.poly_mask([[0, 0, 10, 0, 10, 10, 0, 10], [5, 5, 60, 5, 60, 45, 5, 45]], 40, 50)
    assert two.sum() == 45 * 35 + 100 - 25
    assert IO.poly_mask([[1, 1, 2, 2]], 10, 10).sum() == 0                                 # fewer than 3 vertices
    # a triangle: inside / outside by the even-odd rule on pixel centres, up to the one-pixel boundary band of the 5x walk
    tri = [5.0, 5.0, 30.0, 8.0, 12.0, 33.0]
    mt = IO.poly_mask([tri], 40, 50)
    yy, xx = np.mgrid[0:40, 0:50] + 0.5

    def side(ax, ay, bx, by):
        return (bx - ax) * (yy - ay) - (by - ay) * (xx - ax)

    inside = (side(5, 5, 30, 8) > 0) & (side(30, 8, 12, 33) > 0) & (side(12, 33, 5, 5) > 0)
    assert abs(int(mt.sum()) - int(inside.sum())) <= 40 and (mt.astype(bool) ^ inside).sum() <= 60
    assert IO.poly_mask_resized([box], 40, 50, 100).sum() == round(100 * (100 / 40) * (100 / 50))


def test_polygon_oracle_agrees_with_an_independent_point_in_polygon_test():
    """pycocotools is not installed, so rleFrPoly's restatement is cross-checked against a different algorithm: OpenCV's
    pointPolygonTest at the pixel centres (pixel k spans [k, k+1)).  rleFrPoly walks the boundary on a 5x finer integer grid,
    so the two may disagree only where the centre is closer to the boundary than that grid resolves (measured: 0.21 px)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    total = mismatched = 0
    worst = 0.0
    for _ in range(30):
        h, w = int(rng.integers(32, 80)), int(rng.integers(32, 80))
        n = int(rng.integers(3, 9))
        cx, cy = rng.uniform(0.3, 0.7) * w, rng.uniform(0.3, 0.7) * h
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(0.15, 0.45, n) * min(h, w)               # star-shaped: concave corners included
        xs, ys = cx + rad * np.cos(ang), cy + rad * np.sin(ang)
        mask = IO.poly_mask([np.stack([xs, ys], 1).reshape(-1).tolist()], h, w).astype(bool)
        contour = np.stack([xs, ys], 1).astype(np.float32).reshape(-1, 1, 2)
        dist = np.array([[cv2.pointPolygonTest(contour, (x + 0.5, y + 0.5), True) for x in range(w)] for y in range(h)], np.float32)
        diff = mask != (dist > 0)
        total += h * w
        mismatched += int(diff.sum())
        if diff.any():
            worst = max(worst, float(np.abs(dist[diff]).max()))
        assert mask.sum() > 0
    assert worst < 0.3, worst                     # every disagreement sits on the boundary
    assert mismatched < 2e-3 * total, (mismatched, total)
