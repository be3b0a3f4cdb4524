"""Dump, per kernel of libmaggy_b200_cuda.so, the SASS mnemonics that prove the Blackwell-native path
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA, SYNCS = mbarrier, LDGMC/multimem, .SYS
release/acquire, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents, UBLKCP = cp.async.bulk, UTMAREDG = TMA
reduce). Writes profiles/sass_summary.txt. Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "maggy_b200", "lib", "libmaggy_b200_cuda.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|"
                 r"UTMASTG[A-Z0-9_.]*|UTMAPF[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|LDGMC[A-Z0-9_.]*|"
                 r"STG?[A-Z0-9_.]*MC[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|LDG\.E[A-Z0-9_.]*SYS|STG\.E[A-Z0-9_.]*SYS|"
                 r"REDG?\.E[A-Z0-9_.]*SYS|FENCE[A-Z0-9_.]*|NANOSLEEP[A-Z0-9_.]*|ACQBULK[A-Z0-9_.]*|PREEXIT[A-Z0-9_.]*|UTMAREDG[A-Z0-9_.]*|UTCCP[A-Z0-9_.]*|UTCQMMA[A-Z0-9_.]*)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for tok in PAT.findall(line):
            per[cur][tok] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS evidence per kernel (cuobjdump -sass maggy_b200/lib/libmaggy_b200_cuda.so)", ""]
    for (mangled, cnt), name in zip(per.items(), demangle):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        if not cnt:
            continue
        lines.append(f"{name}")
        lines.append("    " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print(f"wrote {path} ({len(lines)} lines)")


if __name__ == "__main__":
    sys.exit(main())
