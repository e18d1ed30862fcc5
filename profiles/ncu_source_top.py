"""Source-level view of an .ncu-rep captured with --import-source on: the SASS lines with the most warp-stall samples and the
share of executed instructions by kind.  usage: python profiles/ncu_source_top.py file.ncu-rep [n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    inst = [int(r[ix["Instructions Executed"]]) for r in data]
    samp = [int(r[ix["# Samples"]]) for r in data]
    ti, ts = sum(inst), sum(samp)
    print(f"instructions executed (sum over SASS lines) {ti:.4g}, warp-stall samples {ts}")
    kinds = {}
    for r, n in zip(data, inst):
        op = r[ix["Source"]].split()
        op = [t for t in op if not t.startswith("@")]
        k = op[0].split(".")[0] if op else "?"
        kinds[k] = kinds.get(k, 0) + n
    print("executed instructions by opcode:", ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])[:14]))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"{'line':>5} {'SASS':58s} {'samples':>8} {'share':>6} {'executed':>12}  top stall reasons")
    for i in sorted(range(len(data)), key=lambda j: -samp[j])[:top_n]:
        r = data[i]
        st = sorted(((h, int(r[ix[h]])) for h in stall_cols if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {r[ix['Source']].strip()[:58]:58s} {samp[i]:8d} {100 * samp[i] / ts:5.1f}% {inst[i]:12d}  " + ", ".join(f"{h[6:]} {v}" for h, v in st))


if __name__ == "__main__":
    main()
