"""Static evidence from the shipped library, no GPU needed: per kernel the registers / stack (spill) bytes / static shared memory
(`cuobjdump -res-usage`) and the count of the SASS instructions that prove the Blackwell paths (`cuobjdump -sass`; mnemonics from
/opt/skills/guides/B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor
loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

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


def ptxas_spills():
    """mangled kernel name -> (spill store bytes, spill load bytes) from `nvcc -Xptxas -v` over csrc/*.cu (the Makefile's flags)."""
    src = os.path.join(ROOT, "yolopoint_b200", "csrc")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for cu in sorted(glob.glob(os.path.join(src, "*.cu"))):
            cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-Xptxas", "-v",
                   "-c", cu, "-o", os.path.join(tmp, os.path.basename(cu) + ".o")]
            procs.append(subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True, cwd=src))
        for p in procs:
            log = p.communicate()[1]
            for m in re.finditer(r"Function properties for (\S+)\n\s+\d+ bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log):
                out[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return out


def main():
    spills = ptxas_spills()
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
    n_spill = sum(1 for n in names if sum(spills.get(n, (0, 0))))
    print(f"`stack` = bytes of per-thread stack; `spill` = spill store / load bytes per thread reported by `nvcc -Xptxas -v` with the Makefile's flags: "
          f"{len(names) - n_spill} of {len(names)} kernels have none; the tcgen05 conv kernels run under register caps (`__launch_bounds__(256, 2)` = 128, "
          "`__maxnreg__(144)` for the drain kernel so that its 384 threads leave room for the small kernels of other frames) and park 2 - 13 registers' worth of "
          "values on the stack; the opt-in chain kernel spills 700 B.  Where the drain kernel's 15 STL / LDL instructions sit (this build, backward-branch ranges of "
          "`cuobjdump -sass`): 4 stores before the persistent tile loop, the other 11 in the body of the per-TILE loop of the epilogue warps, none inside the "
          "drain-round loop (the LDTM loop) or the TMA / MMA issue loops -- about ten local-memory accesses per thread and output tile.\n")
    print("| kernel | regs | stack B | spill st / ld B | static smem B | instr. | " + " | ".join(MNEMONICS[:9]) + " | FFMA | LDG / STG | ATOMG+RED |")
    print("|---|---:|---:|---:|---:|---:|" + "---:|" * 9 + "---:|---:|---:|")
    for n, s in sorted(zip(names, short), key=lambda t: t[1]):
        r, st, sh, lo = usage[n]
        c = counts[n]
        sp = spills.get(n, (0, 0))
        print(f"| `{s}` | {r} | {st} | {(str(sp[0]) + ' / ' + str(sp[1])) if sum(sp) else ''} | {sh} | {c['_all']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS[:9]) +
              f" | {c['FFMA'] or ''} | {c['LDG']} / {c['STG']} | {(c['ATOMG'] + c['RED']) or ''} |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
