"""Static evidence per kernel from the built library: `python scripts/sass_summary.py > profiles/sass_summary.txt`.

For every kernel of qibo_b200/lib/libqibo_b200.so (cuobjdump -sass / -res-usage): instruction count, registers, stack
(spill) bytes, shared memory, and counts of the instruction classes that matter here -- FP64 (DFMA/DMUL/DADD), packed
FP32 (FFMA2/FMUL2/FADD2), shared-memory (LDS/STS), global (LDG/STG), the TMA / bulk-copy engine (UTMALDG, UTMASTG,
UBLKCP), mbarrier traffic (SYNCS), barriers (BAR), branches, local-memory spills (LDL/STL), tensor-core MMA (none
expected: the per-amplitude work is 2x2 FP64 / FP32 blocks)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qibo_b200", "lib", "libqibo_b200.so")
CLASSES = [
    ("fp64", r"^(DFMA|DMUL|DADD)"), ("fp32x2", r"^(FFMA2|FMUL2|FADD2)"), ("fp32", r"^(FFMA|FMUL|FADD)(?!2)"),
    ("lds", r"^LDS"), ("sts", r"^STS"), ("ldg", r"^(LDG|LD\.)"), ("stg", r"^(STG|ST\.)"), ("tma_ld", r"^UTMALDG"), ("tma_st", r"^UTMASTG"),
    ("bulk", r"^UBLKCP"), ("mbar", r"^SYNCS"), ("bar", r"^BAR"), ("branch", r"^(BRA|BRX|JMP|CALL|RET)"), ("spill", r"^(LDL|STL)"),
    ("mma", r"^(UTC|HMMA|DMMA|IMMA|QMMA|LDTM|STTM)"),
]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED|LOCAL):\d+", line))
            cur = None
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if name and m:
            op = m.group(1)
            kernels[name]["total"] += 1
            for cls, pat in CLASSES:
                if re.match(pat, op):
                    kernels[name][cls] += 1
    nice = demangle(list(kernels))
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a; cuobjdump -sass / -res-usage)")
    cols = ["total"] + [c for c, _ in CLASSES]
    print("kernel | resources | " + " | ".join(cols))
    for k, cnt in kernels.items():
        short = re.sub(r"\(.*", "", nice[k]).replace("qb::", "")
        print(f"{short} | {usage.get(k, '?')} | " + " | ".join(str(cnt.get(c, 0)) for c in cols))
    tot = collections.Counter()
    for cnt in kernels.values():
        tot.update(cnt)
    print("ALL | - | " + " | ".join(str(tot.get(c, 0)) for c in cols))


if __name__ == "__main__":
    main()
