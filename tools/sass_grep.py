"""Count the SASS mnemonics that prove which hardware paths libgpk.so uses (B200_PROFILING.md "What proves a
Blackwell-native kernel").  Runs on the build box (no GPU needed): python tools/sass_grep.py > profiles/rNN_sass_grep.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stheno_b200", "csrc", "libgpk.so")
PATTERNS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "DMMA", "HMMA",
            "IMMA", "LDGSTS", "SYNCS", "UTCBAR", "MUFU.EX2", "MUFU.RSQ64H", "DFMA", "REDG", "ACQBULK"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        for p in PATTERNS:
            if re.search(r"\b" + re.escape(p), line):
                per[cur][p] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} | grep -c <mnemonic>   (sm_100a; {len(per)} kernels)")
    print("total: " + "  ".join(f"{p}={total[p]}" for p in PATTERNS if total[p]))
    print()
    for k, c in per.items():
        if c:
            print(f"{k}: " + "  ".join(f"{p}={c[p]}" for p in PATTERNS if c[p]))


if __name__ == "__main__":
    sys.exit(main())
