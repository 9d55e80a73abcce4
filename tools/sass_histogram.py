"""SASS opcode histogram of the built library -> profiles/r02_sass_histogram.txt (evidence that the hot kernels are
tcgen05 / TMEM / TMA code: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit; HMMA would be the legacy mma.sync path).

    python tools/sass_histogram.py [out_file]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hcflow_b200", "libhcflow_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|STTM(?:\.[A-Za-z0-9_]+)*|UTMALDG(?:\.[A-Z0-9_]+)*|"
                 r"UTMASTG(?:\.[A-Z0-9_]+)*|UBLKCP(?:\.[A-Z0-9_]+)*|UTCBAR(?:\.[A-Z0-9_]+)*|UTCCP(?:\.[A-Z0-9_]+)*|"
                 r"SYNCS(?:\.[A-Z0-9_]+)*|HMMA(?:\.[A-Z0-9_]+)*|HGMMA|QGMMA|IGMMA|FFMA|MUFU(?:\.[A-Z0-9_]+)*|RED(?:\.[A-Z0-9_]+)*)\b")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, check=True).stdout.decode(errors="replace")
    per_fn = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for op in PAT.findall(line):
            per_fn[cur][op.split(".")[0] if op.startswith(("FFMA", "MUFU", "RED", "SYNCS")) else op] += 1
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    with open(out, "w") as f:
        f.write("# cuobjdump -sass hcflow_b200/libhcflow_b200.so | opcode histogram (tools/sass_histogram.py)\n")
        f.write("# whole library\n")
        for op, n in sorted(total.items(), key=lambda kv: (-kv[1], kv[0])):
            f.write("{:8d}  {}\n".format(n, op))
        f.write("# per kernel (tensor-core / TMA opcodes only)\n")
        for fn, c in per_fn.items():
            keep = {k: v for k, v in c.items() if k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "HMMA"))}
            if keep:
                name = subprocess.run(["c++filt", fn], stdout=subprocess.PIPE).stdout.decode().strip() or fn
                f.write("{}\n    {}\n".format(name[:160], ", ".join("{} x{}".format(k, v) for k, v in sorted(keep.items()))))
    print(out)


if __name__ == "__main__":
    main()
