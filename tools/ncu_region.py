"""Stall samples / instructions of source-line regions from an `ncu --page source --csv --print-source cuda,sass` dump.
usage: python tools/ncu_region.py dump.csv file:lo-hi [file:lo-hi ...]"""
import collections
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def load(path):
    rows = list(csv.reader(open(path)))
    cur = hdr = None
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or cur is None:
            continue
        try:
            line = int(r[0])
        except Exception:
            continue
        s, i = num(r[hdr.index("# Samples")]), num(r[hdr.index("Instructions Executed")])
        a = agg.setdefault((cur, line), [0.0, 0.0, collections.Counter(), r[1][:100]])
        a[0] += s; a[1] += i
        a[2].update({hdr[c][6:]: num(r[c]) for c in cols})
        tot += s
    return agg, tot


def main():
    agg, tot = load(sys.argv[1])
    print("total samples", tot)
    for spec in sys.argv[2:]:
        f, rng = spec.split(":")
        lo, hi = (int(v) for v in rng.split("-"))
        ssum = isum = 0.0
        c = collections.Counter()
        for (ff, l), (s, i, st, src) in agg.items():
            if ff == f and lo <= l <= hi:
                ssum += s; isum += i; c.update(st)
                if s > tot * 0.002:
                    print(f"  {100 * s / tot:5.2f}% smp {i / 1e6:7.1f}M inst {l}: {src[:80]} | " + ", ".join(f"{k}:{v:.0f}" for k, v in st.most_common(3)))
        print(f"== {spec}: {100 * ssum / max(tot, 1):.1f}% samples, {isum / 1e6:.0f}M inst; stalls: " +
              ", ".join(f"{k}:{100 * v / max(ssum, 1):.0f}%" for k, v in c.most_common(6)))


if __name__ == "__main__":
    main()
