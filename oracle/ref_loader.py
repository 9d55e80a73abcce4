"""Import the UNMODIFIED reference modules.

TEST / BENCH INFRASTRUCTURE.  /root/reference does not exist on the GPU box: the parity tests there use the
committed fixtures under tests/golden/; the bench's reference arm and the factory test use the unmodified copy that
oracle/build_ref.py stages under oracle/_ref (git-ignored).  The reference's
``utils/util.py`` needs ``natsort`` and ``matplotlib`` at import time;
two three-line stubs under oracle/_stubs/ satisfy that (plus ``lpips`` / ``lmdb`` stand-ins for the
reference's test script and dataset modules, which this image lacks).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("HCFLOW_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "codes", "models", "modules")):
    # GPU box: the unmodified copy staged by oracle/build_ref.py (git-ignored, travels with the snapshot)
    REF_ROOT = os.path.join(_HERE, "_ref")
REF_CODES = os.path.join(REF_ROOT, "codes")
_STUBS = os.path.join(_HERE, "_stubs")


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
