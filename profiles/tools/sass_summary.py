"""SASS opcode summary of the hot kernels of the built library (runs without a GPU):
    python profiles/tools/sass_summary.py [nd np] > profiles/r2_final/sass_opcodes.txt
Counts, per kernel instance of the config-4 (3-D, p=4) and config-3 (p=3) translation units, the
mnemonics that show the Blackwell data path (UBLKCP = TMA bulk copy, SYNCS = mbarrier, LDGSTS =
cp.async), the fp64 pipe (DFMA / DMUL / DADD / MUFU.RCP64H), shared / global / local memory
instructions and the register count.  No UTMALDG / UTC*MMA is expected: the path has no contraction
large enough for tensor cores."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WANT = ["UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "UTCMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG",
        "LDL", "STL", "BAR", "NANOSLEEP"]


def summarise(obj, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
    cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
    sass = subprocess.run(["nvdisasm", cubin], capture_output=True, text=True).stdout.split("\n")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    cur, counts = None, {}
    for ln in sass:
        m = re.match(r"//-+ \.text\.(\S+) -+", ln)
        if m:
            cur = m.group(1) if re.search(pattern, m.group(1)) else None
            if cur:
                counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                op = m.group(1)
                counts[cur]["total"] += 1
                for w in WANT:
                    if op.startswith(w):
                        counts[cur][w] += 1
    for name, c in counts.items():
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"{demangled}\n    registers {regs.get(name, '?')}, {c['total']} SASS instructions: "
              + ", ".join(f"{w} {c[w]}" for w in WANT if c[w]))


if __name__ == "__main__":
    pairs = [(3, 5), (3, 4)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
    for nd, npn in pairs:
        obj = os.path.join(ROOT, "flou.jl_b200", "csrc", "build", f"inst_{nd}_{npn}.o")
        print(f"== {os.path.relpath(obj, ROOT)}  (Euler, Chandrasekhar split form, Cartesian: the benchmark instances)")
        summarise(obj, r"(line_kernel_ws\w*INS_4LCfgILi%dELi%dELi1ELi2ELb1|face_flux_kernelILi%dELi%dELi1ELb1)" % (nd, npn, nd, npn))
