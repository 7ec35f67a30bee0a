"""Opcode histogram per kernel of the shipped libbbmpc.so (cuobjdump -sass): the evidence that the rollout kernels are
tcgen05 / TMEM / bulk-copy code (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR =
tcgen05.commit, SYNCS = mbarrier, USETMAXREG = setmaxnreg).   python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "blackbox_mpc_b200", "libbbmpc.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "USETMAXREG", "UTCATOMSWS", "MUFU", "F2FP", "FFMA2", "FMUL2", "FADD2",
       "STS", "LDS", "LDG", "STG", "REDG", "ATOMS", "MEMBAR", "FENCE", "CCTL", "BAR", "HMMA", "FFMA", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    for name, pretty in zip(kernels, demangle):
        c = kernels[name]
        total = sum(c.values())
        keys = "  ".join(f"{k}={c[k]}" for k in KEY if c[k])
        print(f"{pretty[:110]}\n    instructions={total}  {keys}")


if __name__ == "__main__":
    main()
