"""Static evidence from the shipped library, no GPU needed: per kernel the registers / stack (spill) bytes / static shared memory
(`cuobjdump -res-usage`) and the count of the SASS instructions that prove the Blackwell paths (`cuobjdump -sass`; mnemonics from
/opt/skills/guides/B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor
loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yolopoint_b200", "libyolopoint_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "ATOMG", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = n.replace("yp::", "")
        n = re.sub(r"\(.*$", "", n)          # drop the parameter list
        short.append(n)
    return short


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    arch = set()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    names = sorted(usage)
    short = demangle(names)
    tot = collections.Counter()
    for n in names:
        tot.update(counts[n])
    print("# SASS / resource summary of the shipped `yolopoint_b200/libyolopoint_b200.so` (round 2)\n")
    print("Produced on the build container (no GPU) by `python tools/sass_summary.py` from `cuobjdump -res-usage` and `cuobjdump -sass`;")
    print(f"architectures in the fat binary: {', '.join(sorted(arch)) or 'n/a'}; {len(names)} kernels, {tot['_all']} SASS instructions.\n")
    print("Library totals: " + ", ".join(f"{k} {tot[k]}" for k in MNEMONICS if tot[k]) + ".\n")
    print("`stack` = bytes of per-thread stack (a non-zero value with `local` = 0 is ABI scratch for argument structs / indexed local arrays, "
          "not register spills of the hot loop; the per-kernel spill check is `-Xptxas -v` in `csrc/Makefile`).\n")
    print("| kernel | regs | stack B | static smem B | instr. | " + " | ".join(MNEMONICS[:9]) + " | FFMA | LDG / STG | ATOMG+RED |")
    print("|---|---:|---:|---:|---:|" + "---:|" * 9 + "---:|---:|---:|")
    for n, s in sorted(zip(names, short), key=lambda t: t[1]):
        r, st, sh, lo = usage[n]
        c = counts[n]
        print(f"| `{s}` | {r} | {st} | {sh} | {c['_all']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS[:9]) +
              f" | {c['FFMA'] or ''} | {c['LDG']} / {c['STG']} | {(c['ATOMG'] + c['RED']) or ''} |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
