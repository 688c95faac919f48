"""Ranks source lines of an `ncu --page source --csv --print-source cuda,sass` dump by shared-memory wavefronts.
usage: python tools/ncu_wavefronts.py dump.csv [top-n]"""
import collections
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cur = hdr = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
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
    a = agg[(cur, line)]
    a[0] += num(r[hdr.index("Instructions Executed")])
    a[1] += num(r[hdr.index("L1 Wavefronts Shared")])
    a[2] += num(r[hdr.index("L1 Wavefronts Shared Ideal")])
print(f"total shared wavefronts {sum(a[1] for a in agg.values()) / 1e6:.1f}M, ideal {sum(a[2] for a in agg.values()) / 1e6:.1f}M")
srcs = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    if f not in srcs:
        try:
            srcs[f] = open("audio-flow-rs_b200/csrc/" + f).read().split("\n")
        except Exception:
            srcs[f] = []
    s = srcs[f][l - 1].strip()[:110] if l - 1 < len(srcs[f]) else ""
    print(f"{a[1] / 1e6:6.1f}M wf {a[2] / 1e6:6.1f}M ideal {a[0] / 1e6:6.1f}M inst  {f}:{l}  {s}")
