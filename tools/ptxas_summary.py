"""Summarise the `-Xptxas -v` output of the in-tree build (torchseg_b200/build/*.o.log, written by torchseg_b200.build)
as one table: file, kernel, registers, static smem, spill bytes, stack. Static evidence only (no GPU needed).

Usage: python tools/ptxas_summary.py [> profiles/rNN_ptxas_resources.txt]
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGS = sorted(glob.glob(os.path.join(ROOT, "torchseg_b200", "build", "*.o.log")))


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    out = r.stdout.splitlines() if r.returncode == 0 else names
    short = []
    for s in out:
        s = re.sub(r"\(anonymous namespace\)::", "", s)
        s = re.sub(r"^void ", "", s)
        s = s.split("(")[0]
        short.append(s)
    return short


def main():
    rows = []
    for log in LOGS:
        txt = open(log).read()
        src = os.path.basename(log).replace(".o.log", ".cu")
        for m in re.finditer(
                r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                r"(\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt):
            name, stack, sst, sld, regs, bars, smem = m.groups()
            rows.append((src, name, int(regs), int(smem or 0), int(sst) + int(sld), int(stack)))
    names = demangle([r[1] for r in rows])
    print("%-22s %-58s %5s %8s %6s %6s" % ("source", "kernel", "regs", "smem_B", "spill", "stack"))
    for (src, _n, regs, smem, spill, stack), nm in zip(rows, names):
        print("%-22s %-58s %5d %8d %6d %6d" % (src, nm[:58], regs, smem, spill, stack))
    print("# %d kernels, %d with spills" % (len(rows), sum(1 for r in rows if r[4])))


if __name__ == "__main__":
    main()
