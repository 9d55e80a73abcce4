"""Stub: the reference's utils/util.py imports natsort at module level (util.py:11)."""
natsorted = sorted
natsort = sorted
