"""GPU input pipeline (next-row f4): the sample preparation `train_sam3_lora_native.py:95-172` does on the host with PIL and
pycocotools, bit-exactly, on the GPU.

    prep = GpuPreprocessor(1008)
    x = prep.image(rgb_u8)                      # uint8 [H, W, 3] (numpy or torch, any device) -> fp32 [3, 1008, 1008] on the GPU
    m = prep.rle_masks([(counts, h, w), ...])   # uncompressed or compressed COCO RLE -> bool [N, 1008, 1008] on the GPU

`image` = PILImage.resize((R, R), BILINEAR) + ToTensor + Normalize(0.5, 0.5) (:104-108, :83-86); `rle_masks` =
mask_utils.decode + F.interpolate(mode="nearest") + `> 0.5` (:148-167).  Only the raw uint8 image (3 MB instead of 12 MB of
fp32) and the run lengths cross PCIe.  There is no CPU path: the reference's own host code is the CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's BILINEAR coefficient tables for one axis (host computation in libsam3b): (bounds [out,2], coeffs [out,ks], ks)."""
    lib = L.load()
    ks = lib.sam3b_resample_coeffs(in_size, out_size, None, None)
    if ks <= 0:
        L.check(ks)
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ks), np.int32)
    rc = lib.sam3b_resample_coeffs(in_size, out_size, bounds.ctypes.data_as(C.c_void_p), coeffs.ctypes.data_as(C.c_void_p))
    if rc != ks:
        L.check(rc if rc < 0 else -1)
    return bounds, coeffs, ks


def rle_counts(counts: Union[str, bytes, Sequence[int]]) -> List[int]:
    """COCO RLE counts as a list of run lengths; decodes pycocotools' compressed string form (maskApi.c rleFrString)."""
    if isinstance(counts, bytes):
        counts = counts.decode("ascii")
    if not isinstance(counts, str):
        return [int(c) for c in counts]
    out: List[int] = []
    p, n = 0, len(counts)
    while p < n:
        x, k, more = 0, 0, True
        while more:
            c = ord(counts[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(out) > 2:
            x += out[-2]
        out.append(x)
    return out


class GpuPreprocessor:
    def __init__(self, resolution: int = 1008, mean: float = 0.5, std: float = 0.5, device: Union[str, torch.device] = "cuda"):
        self.resolution, self.mean, self.std = int(resolution), float(mean), float(std)
        self.device = torch.device(device)
        self._tables: Dict[int, Tuple[torch.Tensor, torch.Tensor, int]] = {}

    def _table(self, in_size: int):
        if in_size not in self._tables:
            b, k, ks = resample_coeffs(in_size, self.resolution)
            self._tables[in_size] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), ks)
        return self._tables[in_size]

    def image(self, rgb: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
        if self.device.type != "cuda":
            raise L.Sam3bError("GpuPreprocessor needs a CUDA device (the reference's PIL code is the CPU path)")
        if isinstance(rgb, np.ndarray):
            rgb = np.ascontiguousarray(rgb)
            if not rgb.flags.writeable:          # e.g. np.asarray(PIL image): torch wants a writable buffer
                rgb = rgb.copy()
        t = torch.from_numpy(rgb) if isinstance(rgb, np.ndarray) else rgb
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise L.Sam3bError(f"image: expected uint8 [H, W, 3], got {t.dtype} {tuple(t.shape)}")
        t = t.to(self.device, non_blocking=True).contiguous()
        h, w, R = t.shape[0], t.shape[1], self.resolution
        bx, kx, ksx = self._table(w)
        by, ky, ksy = self._table(h)
        tmp = torch.empty(h, R, 3, device=self.device, dtype=torch.uint8)
        out = torch.empty(3, R, R, device=self.device, dtype=torch.float32)
        L.check(L.load().sam3b_image_resize_normalize(L.ptr(t), h, w, R, L.ptr(bx), L.ptr(kx), ksx, L.ptr(by), L.ptr(ky), ksy,
                                                      L.ptr(tmp), L.ptr(out), self.mean, self.std, L.current_stream()))
        return out

    def rle_masks(self, rles: Sequence[Tuple[Union[str, bytes, Sequence[int]], int, int]]) -> torch.Tensor:
        """rles: (counts, height, width) per mask, `counts` as in a COCO annotation's segmentation dict."""
        if self.device.type != "cuda":
            raise L.Sam3bError("GpuPreprocessor needs a CUDA device (the reference's pycocotools code is the CPU path)")
        R, N = self.resolution, len(rles)
        out = torch.empty(N, R, R, device=self.device, dtype=torch.uint8)
        if N == 0:
            return out.bool()
        cums, offs, hw = [], [0], []
        for counts, h, w in rles:
            c = np.cumsum(np.asarray(rle_counts(counts), dtype=np.int64))
            if len(c) and c[-1] > h * w:
                raise L.Sam3bError(f"RLE covers {int(c[-1])} pixels, more than {h}x{w}")
            cums.append(c.astype(np.uint32))
            offs.append(offs[-1] + len(c))
            hw.append((h, w))
        cum_dev = torch.from_numpy(np.concatenate(cums).view(np.int32)).to(self.device)     # uint32 bit pattern
        offs_dev = torch.tensor(offs, dtype=torch.int32, device=self.device)
        hw_dev = torch.tensor(hw, dtype=torch.int32, device=self.device)
        L.check(L.load().sam3b_rle_masks_nearest(L.ptr(cum_dev), L.ptr(offs_dev), L.ptr(hw_dev), N, R, L.ptr(out), L.current_stream()))
        return out.bool()
