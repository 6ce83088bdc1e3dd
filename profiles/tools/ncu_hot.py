"""Top stall sites of an ncu source page (SASS view):  python ncu_hot.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[idx["# Samples"]]) for r in body)
    print("total samples", tot, "instructions", len(body))
    cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {c: sum(int(r[idx[c]] or 0) for r in body) for c in cols}
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]]))[:n]
    for i in sorted(order):
        r = body[i]
        st = sorted(((int(r[idx[c]] or 0), c) for c in cols), reverse=True)[:2]
        print(f"{i:5d} {int(r[idx['# Samples']]):6d} {100*int(r[idx['# Samples']])/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
