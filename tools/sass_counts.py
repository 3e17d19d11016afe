"""Static SASS profile of selected kernels of libtsb.so: instruction count, memory / fence mnemonics, top opcode groups.
Used for the kernels that have no ncu capture (no GPU needed: `cuobjdump -sass` on the in-tree build).

Usage: python tools/sass_counts.py ce_up p2p_allreduce train_preprocess edge_  [> profiles/rNN_sass_<what>.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torchseg_b200", "lib", "libtsb.so")
MEM = ("LDG", "STG", "LDS", "STS", "RED", "ATOM", "MEMBAR", "FENCE", "ERRBAR", "CCTL", "LD.", "ST.", "LDL", "STL", "UTMA", "UTC", "LDTM")


def main(want):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        if want and not any(w in name for w in want):
            continue
        ops = collections.Counter(m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f))
        rows.append((name, ops))
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("cuobjdump -sass torchseg_b200/lib/libtsb.so — static counts (instructions in the binary, not executed counts)\n")
    for (name, ops), d in zip(rows, dem):
        d = re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0].replace("void ", "")
        grp = collections.Counter()
        for op, c in ops.items():
            grp[op.split(".")[0]] += c
        mem = {k: v for k, v in sorted(ops.items()) if k.startswith(MEM)}
        print("%s: %d instructions" % (d, sum(ops.values())))
        print("    memory / sync: %s" % mem)
        print("    top opcodes:   %s" % grp.most_common(8))


if __name__ == "__main__":
    main(sys.argv[1:])
