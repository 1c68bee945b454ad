"""Writes a seeded COCO-shaped synthetic set: uint8 RGB noise images, 1-4 objects per
image with bbox + one polygon (or uncompressed-RLE) mask each, 1-2 category names.

    python tools/make_synthetic_coco.py data/synthetic_coco --n-train 16 --n-valid 4 --size 1024
"""
import argparse
import json
from pathlib import Path

import numpy as np
from PIL import Image


def make_split(root: Path, split: str, n: int, size: int, rng):
    d = root / split
    d.mkdir(parents=True, exist_ok=True)
    images, anns = [], []
    aid = 1
    for i in range(n):
        arr = rng.integers(0, 256, size=(size, size, 3), dtype=np.uint8)
        name = f"img_{i:05d}.png"
        Image.fromarray(arr).save(d / name)
        images.append({"id": i + 1, "file_name": name, "width": size, "height": size})
        for _ in range(int(rng.integers(1, 5))):
            w, h = (int(v) for v in rng.integers(size // 8, size // 2, size=2))
            x, y = int(rng.integers(0, size - w)), int(rng.integers(0, size - h))
            cat = int(rng.integers(1, 3))
            ann = {"id": aid, "image_id": i + 1, "category_id": cat, "bbox": [x, y, w, h], "area": w * h, "iscrowd": 0}
            if rng.random() < 0.75:
                ann["segmentation"] = [[x, y, x + w, y, x + w, y + h, x + w // 2, y + h // 2, x, y + h]]
            else:  # uncompressed column-major RLE of the box
                counts, pos = [], 0
                for col in range(x, x + w):
                    start = col * size + y
                    counts += [start - pos, h]
                    pos = start + h
                counts.append(size * size - pos)
                ann["segmentation"] = {"size": [size, size], "counts": counts}
            anns.append(ann)
            aid += 1
    coco = {"images": images, "annotations": anns, "categories": [{"id": 1, "name": "crack"}, {"id": 2, "name": "machine part"}]}
    (d / "_annotations.coco.json").write_text(json.dumps(coco))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--n-train", type=int, default=16)
    ap.add_argument("--n-valid", type=int, default=4)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    make_split(Path(a.out), "train", a.n_train, a.size, rng)
    make_split(Path(a.out), "valid", a.n_valid, a.size, rng)
    print(f"wrote {a.n_train}+{a.n_valid} images under {a.out}")


if __name__ == "__main__":
    main()
