"""Import the UNMODIFIED reference modules (authoring container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box; everything that
runs there uses the committed fixtures under tests/golden/ instead.  The reference's
``utils/util.py`` needs ``natsort`` and ``matplotlib`` at import time;
two three-line stubs under oracle/_stubs/ satisfy that.
"""
import os
import sys

REF_ROOT = os.environ.get("HCFLOW_REFERENCE", "/root/reference")
REF_CODES = os.path.join(REF_ROOT, "codes")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available():
    return os.path.isdir(os.path.join(REF_CODES, "models", "modules"))


def load():
    """Returns the reference's ``models.networks`` module (entry: define_G(opt, step))."""
    if not available():
        raise RuntimeError("reference not present at {}".format(REF_ROOT))
    for p in (_STUBS, REF_CODES):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("natsort", "matplotlib"):
        try:
            __import__(name)
        except ImportError:
            pass
    from models import networks  # noqa: E402  (reference package)
    return networks
