"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, integer arithmetic) of the reference's input pipeline (row f4):
`PILImage.resize((1008, 1008), BILINEAR)` + `ToTensor` + `Normalize(0.5, 0.5)` for the image and
`mask_utils.decode` + `F.interpolate(mode="nearest")` + `> 0.5` for RLE masks (train_sam3_lora_native.py:101-108, 146-167).

Third-party arithmetic restated here:
  * Pillow `ImagingResample` (src/libImaging/Resample.c, any Pillow >= 7: 8-bit path with PRECISION_BITS = 22): separable
    triangle filter whose support grows with the down-scale factor, double-precision coefficients rounded to fixed point,
    horizontal pass first with an 8-bit intermediate image, then the vertical pass.  Pinned bit-exactly against the Pillow
    installed in the build container (tests/test_input_oracle.py).
  * pycocotools `rleDecode` (common/maskApi.c): column-major runs, starting with a run of zeros; the compressed string form is
    the LEB128-like code of `rleFrString`.  pycocotools is not installed here ("parity unpinned" for the string decoder; the
    run semantics are pinned by construction against numpy).
  * ATen nearest up/down-sampling index: min(floor(dst * float(in) / out), in - 1) in float32.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle, support 1) filter over the whole axis.
    Returns (bounds [out, 2] = (first source index, tap count), coeffs [out, ksize] int32, ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        w = np.zeros(ksize, np.float64)
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - t if t < 1.0 else 0.0
        ww = w[:xmax].sum() if xmax > 0 else 0.0
        # Pillow accumulates ww sequentially in double; np.sum may pair differently: restate the loop
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            v = w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One resampling pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H, W, C] image, 8-bit result."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)           # [in, other, C]
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(n):
            acc += src[lo + k] * int(kk[o, k])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PILImage.resize((out_w, out_h), BILINEAR) of a uint8 [H, W, C] image: horizontal pass, then vertical pass."""
    h, w = img.shape[:2]
    cur = img
    if w != out_w:
        b, k, _ = resample_coeffs(w, out_w)
        cur = _pass(cur, b, k, axis=1)
    if h != out_h:
        b, k, _ = resample_coeffs(h, out_h)
        cur = _pass(cur, b, k, axis=0)
    return cur


def to_tensor_normalize(img_u8: np.ndarray, mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """ToTensor + Normalize in float32: ((u8 / 255) - mean) / std, [H, W, C] -> [C, H, W]."""
    t = img_u8.astype(np.float32) / np.float32(255.0)
    t = (t - np.float32(mean)) / np.float32(std)
    return np.ascontiguousarray(t.transpose(2, 0, 1))


def rle_decode(counts: Sequence[int], h: int, w: int) -> np.ndarray:
    """pycocotools rleDecode: runs alternate 0,1,0,... over the COLUMN-major pixel order; returns uint8 [h, w]."""
    flat = np.zeros(h * w, np.uint8)
    pos, val = 0, 0
    for c in counts:
        if val:
            flat[pos:pos + c] = 1
        pos += c
        val ^= 1
    return flat.reshape(w, h).T.copy()


def rle_from_string(s: str) -> List[int]:
    """pycocotools rleFrString: LEB128-like, 5 data bits per char (ASCII 48..), sign bit 0x10, counts after the second one are
    stored as differences to the count two positions earlier."""
    cnts: List[int] = []
    p, m = 0, 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if m > 2:
            x += cnts[m - 2]
        cnts.append(x)
        m += 1
    return cnts


def nearest_index(dst: int, in_size: int, out_size: int) -> int:
    """ATen nearest_idx (float32 scale)."""
    if out_size == in_size:
        return dst
    if out_size == 2 * in_size:
        return dst >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    return min(int(np.floor(np.float32(dst) * scale)), in_size - 1)


def rle_mask_resized(counts: Sequence[int], h: int, w: int, out: int) -> np.ndarray:
    """decode + F.interpolate(mask[None, None].float(), (out, out), mode="nearest") > 0.5 -> bool [out, out]."""
    m = rle_decode(counts, h, w)
    ys = np.array([nearest_index(i, h, out) for i in range(out)])
    xs = np.array([nearest_index(i, w, out) for i in range(out)])
    return m[ys][:, xs] > 0


# ------------------------------------------------------------------------------------------------------------------
# Polygon masks: pycocotools `frPyObjects` (list of polygons) -> `merge` -> `decode`, the call sequence of
# train_sam3_lora_native.py:152-156.  Third-party arithmetic restated from the published algorithm of
# pycocotools/common/maskApi.c `rleFrPoly` (pycocotools >= 2.0.6, requirements.txt; not installed here: PARITY UNPINNED
# beyond the known answers in tests/test_input_oracle.py - an axis-aligned integer box polygon covers exactly w*h pixels -
# and an independent cross-check: against OpenCV's point-in-polygon test at the pixel centres, random star-shaped polygons
# differ only at pixels whose centre lies within 0.22 px of the boundary, the quantisation of rleFrPoly's 5x grid).
# ------------------------------------------------------------------------------------------------------------------
def _c_int(v: float) -> int:
    """C's (int) cast: truncation toward zero."""
    return int(v)


def poly_crossings(xy: Sequence[float], h: int, w: int) -> np.ndarray:
    """The sorted column-major positions x*h + y at which rleFrPoly toggles the fill for one polygon (k vertices as
    x0, y0, x1, y1, ...): boundary up-sampled by 5, walked edge by edge along its longer axis, reduced to the points where
    the walk enters a new up-sampled column that maps onto a pixel centre column.  Duplicates are kept (parity counts)."""
    scale = 5.0
    k = len(xy) // 2
    x = [_c_int(scale * float(xy[2 * j]) + 0.5) for j in range(k)]
    y = [_c_int(scale * float(xy[2 * j + 1]) + 0.5) for j in range(k)]
    x.append(x[0])
    y.append(y[0])
    u: List[int] = []
    v: List[int] = []
    for j in range(k):
        xs, xe, ys, ye = x[j], x[j + 1], y[j], y[j + 1]
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx > 0 else 0.0          # dx == dy == 0: C divides 0/0; the single point is (xs, ys) either way
            for d in range(dx + 1):
                t = dx - d if flip else d
                u.append(t + xs)
                v.append(_c_int(ys + s * t + 0.5))
        else:
            s = (xe - xs) / dy
            for d in range(dy + 1):
                t = dy - d if flip else d
                v.append(t + ys)
                u.append(_c_int(xs + s * t + 0.5))
    out: List[int] = []
    for j in range(1, len(u)):
        if u[j] == u[j - 1]:
            continue
        xd = float(u[j] if u[j] < u[j - 1] else u[j] - 1)
        xd = (xd + 0.5) / scale - 0.5
        if math.floor(xd) != xd or xd < 0 or xd > w - 1:
            continue
        yd = float(v[j] if v[j] < v[j - 1] else v[j - 1])
        yd = (yd + 0.5) / scale - 0.5
        yd = 0.0 if yd < 0 else (float(h) if yd > h else yd)
        yd = math.ceil(yd)
        out.append(int(xd) * h + int(yd))
    return np.sort(np.asarray(out, dtype=np.int64))


def poly_mask(polygons: Sequence[Sequence[float]], h: int, w: int) -> np.ndarray:
    """frPyObjects + merge (union) + decode of a COCO polygon list -> uint8 [h, w].  A position's value inside one polygon is
    the parity of the number of crossings at or before it in column-major order (rleFrPoly's run construction merges
    zero-length runs, i.e. equal crossing positions cancel in pairs)."""
    total = np.zeros(h * w, np.uint8)
    for poly in polygons:
        if len(poly) < 6:
            continue
        a = poly_crossings(poly, h, w)
        a = a[a < h * w]
        toggles = np.zeros(h * w + 1, np.int64)
        np.add.at(toggles, a, 1)
        total |= (np.cumsum(toggles[:-1]) & 1).astype(np.uint8)
    return total.reshape(w, h).T.copy()


def poly_mask_resized(polygons: Sequence[Sequence[float]], h: int, w: int, out: int) -> np.ndarray:
    """poly_mask + F.interpolate(mode="nearest") to [out, out] + `> 0.5` (train_sam3_lora_native.py:158-163) -> bool."""
    m = poly_mask(polygons, h, w)
    ys = np.array([nearest_index(i, h, out) for i in range(out)])
    xs = np.array([nearest_index(i, w, out) for i in range(out)])
    return m[ys][:, xs] > 0
