"""Stall samples of an ncu report aggregated by CUDA source line:
    python ncu_lines.py report.ncu-rep object.o kernel-name-substring [N]
The SASS page of the report lists the kernel's instructions in program order; `nvdisasm -g` of the
same object lists them in the same order with `//## File "...", line N` markers (-lineinfo build).
The two are matched by instruction index (and the opcode is cross-checked)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    out, inside, cur = [], False, ("?", 0)
    for ln in text:
        if ln.startswith("//--------------------- .text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(rep, obj, kernel, n=40):
    sass = sass_lines(obj, kernel)
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    hdr = rows[1]
    idx = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    if len(body) != len(sass):
        print(f"warning: {len(body)} instructions in the report, {len(sass)} in the object (different builds?)")
    tot = sum(int(r[idx["# Samples"]]) for r in body)
    cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for (loc, op), r in zip(sass, body):
        a = agg[loc]
        a[0] += int(r[idx["# Samples"]])
        a[1] += int(r[idx["Instructions Executed"]] or 0)
        for c in cols:
            a[2][c] += int(r[idx[c]] or 0)
    src = {}
    print(f"total samples {tot}")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        f, line = loc
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src[f][line - 1].strip()[:90] if 0 < line <= len(src[f]) else ""
        top = ", ".join(f"{k[6:]} {v}" for k, v in a[2].most_common(2))
        print(f"{100 * a[0] / tot:5.1f}%  {a[1]:>12d} inst  {f}:{line:<5d} {text}   [{top}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
