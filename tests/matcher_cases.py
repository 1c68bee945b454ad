"""Case table shared by the matcher golden generator and the matcher tests (tests/golden/make_golden_matcher.py)."""
CASES = {
    # name: (ctor kwargs, B, Q, num_boxes, repeats, repeat_batch, use validity masks)
    "plain": (dict(), 3, 20, [3, 0, 5], 1, 1, False),
    "focal": (dict(focal=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0), 4, 50, [7, 1, 12, 4], 1, 1, False),
    "stable": (dict(focal=True, stable=True, alpha=0.3, gamma=1.5), 2, 32, [6, 9], 1, 1, False),
    "o2m": (dict(focal=True), 2, 40, [4, 6], 3, 1, False),
    "masks": (dict(cost_bbox=5.0), 3, 24, [5, 3, 8], 1, 1, True),
    "more_targets": (dict(), 2, 6, [9, 4], 1, 1, False),
    "aux": (dict(focal=True), 2, 16, [3, 5], 1, 2, False),
    "sam3": (dict(focal=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, alpha=0.25, gamma=2.0), 8, 200, [1, 4, 2, 17, 0, 3, 9, 30], 1, 1, False),
}
