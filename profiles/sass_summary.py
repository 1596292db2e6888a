"""Opcode evidence per kernel of the shipped objects: which kernels use TMA (UTMALDG / UTMASTG / UBLKCP), the FP64 tensor
pipe (DMMA), cp.async (LDGSTS), mbarriers (SYNCS) — and which spill (STL / LDL).  Runs on the CPU box:
    python profiles/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "DFMA", "LDGSTS", "SYNCS", "UTC", "LDTM", "HMMA", "STL", "LDL", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|cmoe::", "", o).split("(")[0] for o in out]


def main():
    print("# SASS opcode histogram per kernel (cuobjdump -sass of cornell-moe_b200/build/*.o, sm_100a)")
    print("# columns: " + " ".join(WATCH) + " | total instructions")
    for obj in sorted(glob.glob(os.path.join(ROOT, "cornell-moe_b200", "build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels = collections.OrderedDict()
        cur = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                op = m.group(1)
                kernels[cur]["_total"] += 1
                for w in WATCH:
                    if op.startswith(w):
                        kernels[cur][w] += 1
        if not kernels:
            continue
        print(f"\n## {os.path.basename(obj)}")
        names = list(kernels)
        for mangled, nice in zip(names, demangle(names)):
            c = kernels[mangled]
            if c["_total"] < 50:
                continue
            cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
            print(f"{nice[:70]:70s} {cols} | {c['_total']}")


if __name__ == "__main__":
    main()
