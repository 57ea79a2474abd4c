"""Per-kernel count of the Blackwell-only SASS instructions in libclipb200.so: UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA bulk tensor load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS
(mbarrier).  Runs anywhere cuobjdump is installed (no GPU needed):

    python profiles/sass_digest.py > profiles/r02_sass_digest.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cli-p_b200", "clipb200", "libclipb200.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTMAPF", "SYNCS", "HMMA", "ELECT", "MEMBAR.ALL.SYS",
             "RED.E.ADD.STRONG.SYS", "LD.E.STRONG.SYS")


def digest(lib=LIB):
    """-> OrderedDict kernel name -> Counter of mnemonics (plus '_instructions')."""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_instructions"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                cur[mn] += 1
    return out


def demangle(names):
    try:
        r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True)
        return dict(zip(names, r.stdout.splitlines()))
    except Exception:
        return {n: n for n in names}


def shorten(name):
    """'void cb::(anonymous namespace)::kern<256, 1, 2>(args...)' -> 'kern<256, 1, 2>'"""
    name = name.replace("cb::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("cb::", "")
    if name.startswith("void "):
        name = name[5:]
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    d = digest(sys.argv[1] if len(sys.argv) > 1 else LIB)
    pretty = demangle(list(d))
    cols = [m for m in MNEMONICS if any(c[m] for c in d.values())]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): Blackwell instruction counts per kernel")
    print("# " + " ".join(f"{c:>9s}" for c in ["instr"] + cols) + "  kernel")
    tot = Counter()
    for name, c in d.items():
        tot.update(c)
        short = shorten(pretty[name])
        print("  " + " ".join(f"{c[k]:9d}" for k in ["_instructions"] + cols) + "  " + short)
    print("# " + " ".join(f"{tot[k]:9d}" for k in ["_instructions"] + cols) + "  TOTAL")


if __name__ == "__main__":
    main()
