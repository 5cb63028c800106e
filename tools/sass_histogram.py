"""SASS opcode histogram of libdatr_b200.so per kernel (runs anywhere: cuobjdump only).  Writes the evidence the
profiling guide asks for -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA), UTCBAR, RED, SYNCS -- to
profiles/.   python tools/sass_histogram.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "datr_b200", "libdatr_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "HMMA", "SYNCS", "RED", "ATOM", "REDG",
       "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "IMAD", "MUFU", "BAR")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcode_histogram.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    total = collections.Counter()
    lines = [f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a), {len(per)} kernels", ""]
    for k, c in per.items():
        for op, n in c.items():
            total[op] += n
        fam = collections.Counter()
        for op, n in c.items():
            for key in KEY:
                if op.split(".")[0] == key or op.startswith(key + "."):
                    fam[op if key.startswith(("UT", "LDTM", "STTM", "RED")) else key] += n
        lines.append(f"{k[:140]}  [{sum(c.values())} instructions]")
        lines.append("    " + ", ".join(f"{op} {n}" for op, n in sorted(fam.items(), key=lambda kv: -kv[1])))
    lines += ["", "## Blackwell-native instructions over the whole library"]
    for op, n in sorted(total.items(), key=lambda kv: -kv[1]):
        if op.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "RED", "HMMA")):
            lines.append(f"  {op:40s} {n}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-30:]))


if __name__ == "__main__":
    main()
