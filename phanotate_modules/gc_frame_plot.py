"""Argmax / argmin frame helpers of the boundary (reference gc_frame_plot.py:7-28); the sliding
window itself runs on the GPU (csrc/pipeline.cuh: scan_range)."""


def max_idx(a, b, c):
    if a > b:
        return 1 if a > c else 3
    return 2 if b > c else 3


def min_idx(a, b, c):
    if a > b:
        return 3 if b > c else 2
    return 3 if a > c else 1
