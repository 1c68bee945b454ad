"""Host logic of the train_sam3_lora_native surface: YAML schema, COCO dataset (polygon + RLE), shapes."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

ROOT = Path(__file__).resolve().parents[1]


def test_shipped_configs_have_the_keys_the_native_cli_reads():
    for name in ("full_lora_config.yaml", "bench_r16_config.yaml", "minimal_lora_config.yaml", "crack_detection_config.yaml"):
        cfg = yaml.safe_load((ROOT / "configs" / name).read_text())
        assert {"rank", "alpha", "dropout", "target_modules"} <= set(cfg["lora"])
        assert {"learning_rate", "weight_decay", "data_dir", "batch_size", "num_epochs"} <= set(cfg["training"])
        assert "output_dir" in cfg["output"]
    full = yaml.safe_load((ROOT / "configs" / "full_lora_config.yaml").read_text())
    assert (full["lora"]["rank"], full["lora"]["alpha"], full["training"]["batch_size"]) == (32, 64, 8)


def test_synthetic_coco_and_dataset(tmp_path):
    out = tmp_path / "coco"
    subprocess.run([sys.executable, str(ROOT / "tools" / "make_synthetic_coco.py"), str(out), "--n-train", "3", "--n-valid", "1",
                    "--size", "128"], check=True)
    from sam3_lora_b200.train_native import COCOSegmentDataset, _ann_to_mask, collate

    ds = COCOSegmentDataset(out, "train")
    assert len(ds) == 3
    item = ds[0]
    assert item["image"].shape == (3, 1008, 1008) and item["image"].dtype == torch.float32
    assert -1.0 <= item["image"].min() and item["image"].max() <= 1.0
    assert item["mask"].shape == (1, 72, 72) and 0 < item["mask"].sum() < 72 * 72
    assert isinstance(item["prompt"], str)
    b = collate([ds[0], ds[1]])
    assert b["image"].shape == (2, 3, 1008, 1008) and b["mask"].shape == (2, 1, 72, 72) and len(b["prompt"]) == 2
    # RLE and polygon of the same box agree
    coco = json.loads((out / "train" / "_annotations.coco.json").read_text())
    x, y, w, h = 10, 20, 30, 40
    counts, pos = [], 0
    for col in range(x, x + w):
        start = col * 128 + y
        counts += [start - pos, h]
        pos = start + h
    counts.append(128 * 128 - pos)
    m_rle = _ann_to_mask({"segmentation": {"size": [128, 128], "counts": counts}}, 128, 128)
    assert m_rle.sum() == w * h and m_rle[y:y + h, x:x + w].all()
    with pytest.raises(FileNotFoundError):
        COCOSegmentDataset(out, "test")


def test_trainer_refuses_to_run_without_gpu(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sam3_lora_b200.train_native import SAM3TrainerNative

    with pytest.raises(RuntimeError, match="CUDA"):
        SAM3TrainerNative(str(ROOT / "configs" / "minimal_lora_config.yaml"))


def test_dataset_items_become_reference_datapoints_and_collate(tmp_path):
    """with_instances=True: per-object normalised xyxy boxes + [R, R] boolean segments, turned into the reference's own
    Datapoint / BatchedDatapoint by its collator (needs the reference under baseline/_ref; skipped otherwise)."""
    from sam3_lora_b200 import sam3_bridge

    if sam3_bridge.reference_root() is None:
        pytest.skip("reference not installed under baseline/_ref")
    out = tmp_path / "coco"
    subprocess.run([sys.executable, str(ROOT / "tools" / "make_synthetic_coco.py"), str(out), "--n-train", "2", "--n-valid", "1",
                    "--size", "96"], check=True)
    from sam3_lora_b200.train_native import COCOSegmentDataset, collate_sam3

    ds = COCOSegmentDataset(out, "train", with_instances=True)
    a = ds[0]
    n = a["boxes"].shape[0]
    assert n >= 1 and a["segments"].shape == (n, 1008, 1008) and a["segments"].dtype == torch.bool
    assert (a["boxes"] >= 0).all() and (a["boxes"] <= 1).all() and (a["boxes"][:, 2:] > a["boxes"][:, :2]).all()
    # the box of an object encloses its mask (both were scaled to the model resolution)
    ys, xs = a["segments"][0].nonzero(as_tuple=True)
    x0, y0, x1, y1 = (a["boxes"][0] * 1008).tolist()
    assert xs.min() >= x0 - 11 and xs.max() <= x1 + 11 and ys.min() >= y0 - 11 and ys.max() <= y1 + 11
    batch = collate_sam3([ds[0], ds[1]])
    assert type(batch).__name__ == "BatchedDatapoint" and batch.img_batch.shape == (2, 3, 1008, 1008)
    assert len(batch.find_text_batch) >= 1 and all(isinstance(t, str) for t in batch.find_text_batch)
    tgt = batch.find_targets[0]
    assert int(tgt.num_boxes.sum()) == n + ds[1]["boxes"].shape[0]


def test_sam3_objective_refuses_random_base_weights(tmp_path, monkeypatch):
    """adapters trained on a random base are useless - no checkpoint means an error unless explicitly allowed."""
    import sam3_lora_b200.train_native as T

    cfg = yaml.safe_load((ROOT / "configs" / "full_lora_config.yaml").read_text())       # ships without allow_random_init
    assert not (cfg.get("model") or {}).get("allow_random_init", False)
    cfg["output"]["output_dir"] = str(tmp_path / "o")
    p = tmp_path / "c.yaml"
    p.write_text(yaml.safe_dump(cfg))
    monkeypatch.delenv("SAM3_CHECKPOINT", raising=False)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    with pytest.raises(RuntimeError, match="checkpoint"):
        T.SAM3TrainerNative(str(p))


def test_objective_normalization_is_validated_and_passed_through():
    """sam3_loss.py:53: 'global' | 'local' | 'none'; anything else is refused before the reference is imported."""
    import pytest

    from sam3_lora_b200 import sam3_bridge, sam3_step

    with pytest.raises(ValueError):
        sam3_step.build_objective(native=False, normalization="per_rank")
    try:
        sam3_bridge.import_reference()
    except Exception:   # noqa: BLE001  (reference not installed under baseline/_ref: nothing more to check here)
        pytest.skip("reference not importable")
    _, wrapper = sam3_step.build_objective(native=False, normalization="global")
    assert wrapper.normalization == "global"
    _, wrapper = sam3_step.build_objective(native=False)
    assert wrapper.normalization == "local"      # train_sam3_lora_native.py:791
