"""Static report of the built library (no GPU needed): per kernel the registers / shared memory / spills ptxas assigned
(cuobjdump --dump-resource-usage) and the count of the SASS mnemonics that show which hardware path a kernel uses
(UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier, FFMA2 = packed fp32, ACQBULK / PDL:
griddepcontrol).  Usage: python scripts/static_report.py > profiles/<round>_static_resources_sass.txt"""
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fem-simulation_b200", "libfemb200.so")
MNEMONICS = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "FFMA2", "FFMA", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "ACQBULK",
             "LDGDEPBAR", "SHFL", "MUFU", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+.*)", res):
        usage[m.group(1)] = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), dict.fromkeys(MNEMONICS, 0))
            cur["_n"] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["_n"] += 1
            op = m.group(1)
            if op in cur:
                cur[op] += 1
    names = demangle(sorted(counts))
    short = lambda s: re.sub(r"\(.*", "", s).replace("femb::", "").replace("(anonymous namespace)::", "")
    print(f"# {os.path.relpath(LIB)}: {len(counts)} kernels (sm_100a); columns: registers, static shared bytes, local (spill) bytes, SASS")
    print(f"# instructions, then the count of selected mnemonics (zero counts omitted)")
    for k in sorted(counts, key=lambda k: short(names[k])):
        u = usage.get(k, {})
        c = counts[k]
        picks = " ".join(f"{op}={c[op]}" for op in MNEMONICS if c[op])
        print(f"{short(names[k]):<58} reg={u.get('REG', '?'):>3} smem={u.get('SHARED', '?'):>6} local={u.get('LOCAL', '?'):>4} "
              f"sass={c['_n']:>6}  {picks}")


if __name__ == "__main__":
    sys.exit(main())
