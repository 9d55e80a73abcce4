"""Stub: the reference's dataset classes import lmdb at module level (data/GTLQ_dataset.py:4); image-folder datasets
never call it."""


def open(*args, **kwargs):
    raise RuntimeError("lmdb is not installed in this image (stub)")
