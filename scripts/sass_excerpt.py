"""SASS evidence of TMA / mbarrier use: python scripts/sass_excerpt.py > profiles/r02_sass_tma_mbarrier.txt
Static counts per kernel template from `cuobjdump -sass libbellman.so`, one representative instance each."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "optimal-control-dynamic-programming_b200", "libbellman.so")
PAT = re.compile(r"\b(UTMALDG\.\dD|UTMAPF\.L2\.\dD|SYNCS\.[A-Z0-9.]+|LDC\.64)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangled = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                               capture_output=True, text=True).stdout.splitlines()
    names = iter(demangled)
    kernels = collections.OrderedDict()     # template name -> list of (full name, Counter, sample lines)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = next(names, m.group(1))
            mm = re.search(r"(k_\w+)(<.*>)?\(", full)
            short = (mm.group(1) + re.sub(r"\((int|bool)\)", "", mm.group(2) or "")) if mm else full
            base = mm.group(1) if mm else full
            cur = (short, collections.Counter(), {})
            kernels.setdefault(base, []).append(cur)
            continue
        if cur is None:
            continue
        m = PAT.search(line)
        if m and "/*" in line:
            cur[1][m.group(1)] += 1
            cur[2].setdefault(m.group(1), re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
    print("# SASS evidence of TMA (UTMALDG / UTMAPF) and mbarrier (SYNCS.*) use in libbellman.so, sm_100a, round 2.")
    print("# Made by scripts/sass_excerpt.py from `cuobjdump -sass libbellman.so`: static counts per kernel template,")
    print("# one representative instance each with one sample line per mnemonic (LDC.64 = constant-bank table reads).")
    for base, insts in kernels.items():
        best = max(insts, key=lambda t: sum(v for k, v in t[1].items() if k != "LDC.64"))
        tma = {k: v for k, v in best[1].items() if k != "LDC.64" or base == "k_stage_wide"}
        if not any(k.startswith(("UTMA", "SYNCS")) for k in tma):
            continue
        print("\n== %s   (%d template instance%s; shown: %s)" % (base, len(insts), "" if len(insts) == 1 else "s", best[0]))
        print("   " + ", ".join("%s x%d" % kv for kv in sorted(tma.items())))
        for k in sorted(tma):
            print("   " + best[2][k])


if __name__ == "__main__":
    main()
