"""Stub: the reference's utils/util.py imports matplotlib at module level (util.py:36-38)."""


def use(*args, **kwargs):
    return None
