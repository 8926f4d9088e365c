"""Per-kernel SASS opcode histogram of the built library (evidence for the tcgen05 / TMA claims):
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt
Counts, per kernel of vit-grid-model_b200/libvitgrid.so, the tensor-core and TMA mnemonics B200_PROFILING.md names
(UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-store,
UBLKCP = bulk copy,
HMMA = legacy mma.sync) and the ten most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vit-grid-model_b200", "libvitgrid.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "HMMA", "SYNCS", "MUFU", "LDGSTS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCQMMA"):
                cur["~" + full] += 1
    total = collections.Counter()
    demangle = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines() if per else []
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels")
    for (name, cnt), pretty in zip(per.items(), demangle or per.keys()):
        keyed = {k: v for k, v in cnt.items() if k in KEY or k.startswith("~")}
        for k, v in cnt.items():
            if k in KEY:
                total[k] += v
        if not any(k in cnt for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "LDTM")) and "-v" not in sys.argv:
            continue
        n = sum(v for k, v in cnt.items() if not k.startswith("~"))
        short = re.sub(r"\(.*", "", pretty)[:110]
        print(f"\n{short}   [{n} instructions]")
        print("   " + "  ".join(f"{k}={v}" for k, v in sorted(keyed.items())))
        print("   top: " + "  ".join(f"{k}={v}" for k, v in cnt.most_common(14) if not k.startswith("~")))
    print("\n# library totals: " + "  ".join(f"{k}={total[k]}" for k in KEY if total[k]))


if __name__ == "__main__":
    main()
