"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per top-level FUNCTION of af_fused.cu
(inlined code is attributed to the function whose source lines it came from).
usage: python tools/ncu_regions.py dump.csv [git-rev-of-sources|WORKTREE]"""
import collections, csv, re, subprocess, sys

def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0

def main():
    path = sys.argv[1]; rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    rows = list(csv.reader(open(path)))
    cur = hdr = None
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    stall_cols = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No":
            hdr = r
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or cur is None: continue
        try: line = int(r[0])
        except Exception: continue
        a = agg[(cur, line)]
        a["inst"] += num(r[hdr.index("Instructions Executed")]); a["smp"] += num(r[hdr.index("# Samples")])
        for i, h in stall_cols: a[h] += num(r[i])
    if rev == "WORKTREE": src = open("audio-flow-rs_b200/csrc/af_fused.cu").read().split("\n")
    else: src = subprocess.run(["git", "show", f"{rev}:audio-flow-rs_b200/csrc/af_fused.cu"], capture_output=True, text=True).stdout.split("\n")
    marks = []
    pat = re.compile(r"^(?:__device__|__global__|static|template|size_t|cudaError_t).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, l in enumerate(src, 1):
        if l.startswith("template"): continue
        m = pat.match(l)
        if m and not l.startswith(" "): marks.append((i, m.group(1)))
    def region(f, line):
        if f != "af_fused.cu": return f
        name = "pre"
        for i, n in marks:
            if line >= i: name = n
        return name
    reg = collections.defaultdict(lambda: collections.defaultdict(float))
    for (f, l), v in agg.items():
        r = reg[region(f, l)]
        for k, x in v.items(): r[k] += x
    tot = sum(v["inst"] for v in reg.values()); ts = sum(v["smp"] for v in reg.values())
    print(f"total warp instructions {tot:.3e}, samples {ts:.0f}")
    for k, v in sorted(reg.items(), key=lambda kv: -kv[1]["smp"]):
        top = sorted(((x, h) for h, x in v.items() if h.startswith("stall_")), reverse=True)[:3]
        tops = ", ".join(f"{h[6:]} {100*x/max(ts,1):.1f}%" for x, h in top if x > 0)
        print(f"{100*v['inst']/tot:5.1f}% inst {100*v['smp']/max(ts,1):5.1f}% smp  {k:22s} top stalls: {tops}")

if __name__ == "__main__":
    main()
