#!/bin/bash
# Turns the .ncu-rep files of tools/gpu_profile_final.sh (gpurun_out/) into the tracked summaries under profiles/.
# usage: tools/make_profiles.sh <tag>      e.g. r01_v7
set -e
tag=$1
for k in fused fused_vad scan; do
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > /tmp/raw_$k.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/raw_$k.csv > profiles/${tag}_${k}_ncu_summary.txt
done
ncu -i gpurun_out/prof_fused.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_fused.csv 2>/dev/null
python tools/ncu_regions.py /tmp/src_fused.csv WORKTREE > profiles/${tag}_fused_by_function.txt
python tools/ncu_lines.py /tmp/src_fused.csv HEAD 40 > profiles/${tag}_fused_hot_lines.txt 2>/dev/null || true
cp gpurun_out/launches.csv profiles/${tag}_launches.csv
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
rows = list(csv.reader(open('/tmp/raw_fused.csv')))
hdr, vals = rows[0], rows[2]
g = lambda k: float(vals[hdr.index(k)].replace(',', ''))
unit = lambda k: rows[1][hdr.index(k)]
def to_bytes(k):
    v, u = g(k), unit(k)
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
json.dump({"af_fused_kernel_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": f"profiles/{tag}_fused_ncu_summary.txt (ncu --set full, cfg2 VAD off, one launch)",
           "algorithmic_bytes_per_launch": 2211840000}, open('profiles/traffic.json', 'w'), indent=1)
print(open('profiles/traffic.json').read())
PY
