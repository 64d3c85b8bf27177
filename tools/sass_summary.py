"""Opcode evidence per kernel: `cuobjdump -sass` of libcwm_b200.so, counting the SASS mnemonics that prove the Blackwell
paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP,
mma.sync -> HMMA).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "counterfactualworldmodels_b200", "libcwm_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "FFMA2",
        "LDGSTS", "SYNCS", "BAR.SYNC"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            counts[cur] = collections.Counter()
            counts[cur]["_instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_instructions"] += 1
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[cur][k] += 1
            elif op.startswith(k):
                counts[cur][k] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes): opcode counts per kernel")
    print(f"# columns: instructions | " + " | ".join(KEYS))
    total = collections.Counter()
    for fn, c in counts.items():
        short = fn[:fn.index(">(") + 1] if ">(" in fn else re.sub(r"\(.*", "", fn)
        short = short.replace("(int)", "").replace("(bool)", "").replace("void ", "")
        row = [str(c["_instructions"])] + [str(c[k]) for k in KEYS]
        print(f"{short[:110]:110s} " + " ".join(f"{v:>6s}" for v in row))
        total.update(c)
    print(f"{'TOTAL':110s} " + " ".join(f"{str(total[k]):>6s}" for k in ["_instructions"] + KEYS))


if __name__ == "__main__":
    sys.exit(main())
