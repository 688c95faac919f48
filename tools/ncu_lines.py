"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per source line.
usage: python tools/ncu_lines.py dump.csv [git-rev-of-sources] [top-n]"""
import collections
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    path = sys.argv[1]
    rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    cur, hdr = None, None
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    tot = totsamp = 0.0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur is None:
            continue
        try:
            line = int(r[0])
        except Exception:
            continue
        inst = num(r[hdr.index("Instructions Executed")])
        samp = num(r[hdr.index("# Samples")])
        wf = num(r[hdr.index("L1 Wavefronts Shared")])
        wfi = num(r[hdr.index("L1 Wavefronts Shared Ideal")])
        a = agg[(cur, line)]
        a[0] += inst; a[1] += samp; a[2] += wf; a[3] += wfi
        tot += inst; totsamp += samp
    srcs = {}
    for f in {k[0] for k in agg}:
        out = subprocess.run(["git", "show", f"{rev}:audio-flow-rs_b200/csrc/{f}"], capture_output=True, text=True).stdout
        srcs[f] = out.split("\n")
    print(f"total warp instructions {tot:.3e}, samples {totsamp:.0f}")
    for (f, l), (inst, samp, wf, wfi) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        src = srcs[f][l - 1].strip()[:100] if l - 1 < len(srcs.get(f, [])) else ""
        print(f"{100 * inst / tot:5.1f}% inst {100 * samp / max(totsamp, 1):5.1f}% smp {wf / 1e6:7.1f}M wf ({wfi / 1e6:6.1f}M ideal)  {f}:{l}  {src}")


if __name__ == "__main__":
    main()
