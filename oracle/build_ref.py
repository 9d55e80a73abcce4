"""Stage the UNMODIFIED reference under oracle/_ref (git-ignored; travels to the GPU box with the snapshot).

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python (no build system, no setup.py / pyproject, so it can be
neither compiled nor pip-installed); "building" it means copying its source tree as it lies under /root/reference,
byte for byte, into oracle/_ref/codes.  Nothing under oracle/_ref is tracked by git and nothing of the product imports
it: the only users are ``bench.py --impl reference`` / ``bench.py``'s ``gpu_eager_baseline`` leg (which time the
reference's own ``networks.define_G`` net on the GPU box's host cores and, through stock PyTorch / cuDNN, on the B200)
and the tests that run the reference's own factory.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: uses the staged copy)
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("HCFLOW_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
SUBDIRS = ["codes/models", "codes/utils", "codes/options", "codes/data"]
FILES = ["codes/test_HCFlow.py", "codes/train_HCFlow.py"]     # the reference's entry scripts (tests/ref_script_worker.py)


def build(verbose=False):
    """Returns the staged root (oracle/_ref) or None when there is neither a reference nor a staged copy."""
    if not os.path.isdir(os.path.join(SRC, "codes", "models", "modules")):
        return DST if os.path.isdir(os.path.join(DST, "codes", "models", "modules")) else None
    for sub in SUBDIRS:
        s, d = os.path.join(SRC, sub), os.path.join(DST, sub)
        if not os.path.isdir(s):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        cmp = filecmp.dircmp(s, d, ignore=["__pycache__"])
        assert not cmp.diff_files and not cmp.left_only, (sub, cmp.diff_files, cmp.left_only)
    for rel in FILES:
        s = os.path.join(SRC, rel)
        if os.path.isfile(s):
            shutil.copyfile(s, os.path.join(DST, rel))
            assert filecmp.cmp(s, os.path.join(DST, rel), shallow=False), rel
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of {}/codes/{{models,utils,options,data,test_HCFlow.py,train_HCFlow.py}} staged by oracle/build_ref.py.\n"
                "Not tracked by git, not imported by the product.\n".format(SRC))
    if verbose:
        print("staged", DST)
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
