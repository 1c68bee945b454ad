"""GPU input pipeline (next-row f4): the sample preparation `train_sam3_lora_native.py:95-172` does on the host with PIL and
pycocotools, bit-exactly, on the GPU.

    prep = GpuPreprocessor(1008)
    x = prep.image(rgb_u8)                      # uint8 [H, W, 3] (numpy or torch, any device) -> fp32 [3, 1008, 1008] on the GPU
    m = prep.rle_masks([(counts, h, w), ...])   # uncompressed or compressed COCO RLE -> bool [N, 1008, 1008] on the GPU

    m = prep.polygon_masks([(polys, h, w), ...])  # COCO polygon lists -> bool [N, 1008, 1008] on the GPU
    u8 = prep.decode_jpeg(jpeg_bytes)           # nvJPEG (torchvision.io.decode_jpeg(device=cuda)) -> uint8 [H, W, 3] on the GPU

`image` = PILImage.resize((R, R), BILINEAR) + ToTensor + Normalize(0.5, 0.5) (:104-108, :83-86); `rle_masks` =
mask_utils.decode + F.interpolate(mode="nearest") + `> 0.5` (:148-167); `polygon_masks` = mask_utils.frPyObjects + merge + decode
(pycocotools rleFrPoly) + the same resize (:152-163).  Only the raw uint8 image (3 MB instead of 12 MB of fp32), the run
lengths and the polygon vertices cross PCIe.  There is no CPU path: the reference's own host code is the CPU path.
`Prefetcher` overlaps the host side of the NEXT batch (file reads, annotation parsing, H2D copies from pinned memory on a
side stream) with the current step, which the reference's `num_workers=0` loader (:831) does not.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

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


    # ---- polygons ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _polygon_edges(objs: Sequence[Tuple[Sequence[Sequence[float]], int, int]]):
        """Host part of rleFrPoly: (int)(5 * coord + .5) per vertex and the per-edge point counts (O(vertices))."""
        edges, counts, list_ofs = [], [], [0]
        n_lists = 0
        for polys, h, w in objs:
            for poly in polys:
                if len(poly) < 6:
                    continue
                xy = np.asarray(poly, dtype=np.float64).reshape(-1, 2)
                pts = np.trunc(5.0 * xy + 0.5).astype(np.int64)              # C's (int) cast truncates toward zero
                nxt = np.roll(pts, -1, axis=0)
                k = pts.shape[0]
                e = np.empty((k, 8), np.int32)
                e[:, 0], e[:, 1], e[:, 2], e[:, 3] = pts[:, 0], pts[:, 1], nxt[:, 0], nxt[:, 1]
                e[:, 4], e[:, 5], e[:, 6], e[:, 7] = n_lists, h, w, 0
                e[0, 7] = 1
                edges.append(e)
                counts.append(np.maximum(np.abs(nxt[:, 0] - pts[:, 0]), np.abs(nxt[:, 1] - pts[:, 1])) + 1)
                n_lists += 1
            list_ofs.append(n_lists)
        if not edges:
            return None, None, list_ofs
        cnt = np.concatenate(counts)
        pt_start = np.zeros(cnt.shape[0] + 1, np.int64)
        np.cumsum(cnt, out=pt_start[1:])
        if pt_start[-1] >= 2 ** 31:
            raise L.Sam3bError("polygon_masks: boundary too long for 32-bit point offsets")
        return np.concatenate(edges), pt_start.astype(np.int32), list_ofs

    def polygon_masks(self, objs: Sequence[Tuple[Sequence[Sequence[float]], int, int]]) -> torch.Tensor:
        """objs: (polygons, height, width) per object, `polygons` = a COCO annotation's segmentation list
        [[x0, y0, x1, y1, ...], ...].  Returns bool [N, R, R] on the GPU."""
        if self.device.type != "cuda":
            raise L.Sam3bError("GpuPreprocessor needs a CUDA device (the reference's pycocotools code is the CPU path)")
        R, N = self.resolution, len(objs)
        out = torch.empty(N, R, R, device=self.device, dtype=torch.uint8)
        if N == 0:
            return out.bool()
        edges, pt_start, list_ofs = self._polygon_edges(objs)
        lib = L.load()
        hw = torch.tensor([(h, w) for _, h, w in objs], dtype=torch.int32, device=self.device)
        lofs = torch.tensor(list_ofs, dtype=torch.int32, device=self.device)
        if edges is None:
            keys, n_keys = None, 0
        else:
            e_dev = torch.from_numpy(edges).to(self.device)
            p_dev = torch.from_numpy(pt_start).to(self.device)
            total = int(pt_start[-1])
            keys = torch.empty(total, dtype=torch.int64, device=self.device)
            L.check(lib.sam3b_poly_crossings(L.ptr(e_dev), L.ptr(p_dev), edges.shape[0], total, L.ptr(keys), L.current_stream()))
            keys = torch.sort(keys).values           # toggles of list l, ascending, then the INT64_MAX fillers
            n_keys = total
        L.check(lib.sam3b_poly_masks_nearest(L.ptr(keys), n_keys, L.ptr(lofs), L.ptr(hw), N, R, L.ptr(out), L.current_stream()))
        return out.bool()

    # ---- JPEG -------------------------------------------------------------------------------------------------------
    def decode_jpeg(self, data: Union[bytes, bytearray, np.ndarray, torch.Tensor]) -> torch.Tensor:
        """JPEG bytes -> uint8 [H, W, 3] on the GPU through nvJPEG (torchvision.io.decode_jpeg(device=cuda)); feed it to
        `image`.  nvJPEG's inverse DCT / chroma up-sampling differ from libjpeg-turbo's (what PIL uses) by a few grey levels,
        so this leg is close to, not bit-identical with, `PILImage.open(...).convert("RGB")`; PNG / BMP files stay on PIL."""
        if self.device.type != "cuda":
            raise L.Sam3bError("GpuPreprocessor needs a CUDA device")
        from torchvision.io import ImageReadMode, decode_jpeg  # noqa: PLC0415

        if isinstance(data, (bytes, bytearray)):
            data = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        elif isinstance(data, np.ndarray):
            data = torch.from_numpy(np.ascontiguousarray(data))
        chw = decode_jpeg(data, mode=ImageReadMode.RGB, device=self.device)
        return chw.permute(1, 2, 0).contiguous()


def _record_stream(obj, stream) -> None:
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for x in obj:
            _record_stream(x, stream)
    elif isinstance(obj, dict):
        for x in obj.values():
            _record_stream(x, stream)
    elif hasattr(obj, "__dataclass_fields__"):
        for f in obj.__dataclass_fields__:
            _record_stream(getattr(obj, f), stream)


class Prefetcher:
    """Background thread that prepares the next items of an iterable (file reads, JSON / polygon parsing, pinned staging and
    the asynchronous H2D copies they issue on `stream`) while the training step of the current batch runs.  The reference's
    loader is synchronous (`num_workers=0`, train_sam3_lora_native.py:831): its step waits for PIL + pycocotools every time.

        for batch in Prefetcher(loader, depth=2, device=dev): ...

    Items are produced in order; an exception in the worker is re-raised at the consumer.  CUDA work issued by the producer
    goes to a private stream; the consumer's current stream waits for the event recorded after each item."""

    def __init__(self, iterable: Iterable, depth: int = 2, device: Optional[Union[str, torch.device]] = None,
                 transform: Optional[Callable] = None):
        self.iterable, self.depth, self.transform = iterable, max(1, int(depth)), transform
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(self.device) if self.device is not None and self.device.type == "cuda" else None

    def __len__(self):
        return len(self.iterable)

    def __iter__(self) -> Iterator:
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        _END, _ERR = object(), object()

        def produce():
            for item in self.iterable:               # dataset / collate code runs here: its GPU work lands on self.stream
                if stop.is_set():
                    return
                if self.transform is not None:
                    item = self.transform(item)
                ev = None
                if self.stream is not None:
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                q.put((item, ev))
            q.put((_END, None))

        def work():
            try:
                if self.stream is not None:
                    torch.cuda.set_device(self.device)
                    with torch.cuda.stream(self.stream):
                        produce()
                else:
                    produce()
            except BaseException as e:  # noqa: BLE001 - handed to the consumer
                q.put((_ERR, e))

        t = threading.Thread(target=work, daemon=True, name="sam3b-prefetch")
        t.start()
        try:
            while True:
                item, ev = q.get()
                if item is _END:
                    return
                if item is _ERR:
                    raise ev
                if ev is not None:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(ev)
                    _record_stream(item, cur)        # allocated on the producer's stream, consumed on this one
                yield item
        finally:
            stop.set()
            while t.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)
