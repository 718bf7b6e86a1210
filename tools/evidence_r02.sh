#!/bin/bash
# Round-2 evidence run (one B200): per-config bench lines, ncu launch list of the timed sweep of bench.py, ncu --set full summaries.
# Usage (through gpurun): bash tools/evidence_r02.sh
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
# ---- per-config lines (BASELINE configs 1, 2, 3 on the 2x2 cell; config 5 shape on the 4x4 cell at N=1)
timeout 300 python bench.py --D 2 --chi 20 --nx 2 --ny 2 --steps 40 --warmup 10 > gpurun_out/r02_bench_config1_D2_chi20.json 2>/dev/null
timeout 300 python bench.py --D 4 --chi 64 --nx 2 --ny 2 --steps 40 --warmup 10 > gpurun_out/r02_bench_config2_D4_chi64.json 2>/dev/null
timeout 400 python bench.py --D 6 --chi 144 --nx 2 --ny 2 --steps 10 --warmup 4 > gpurun_out/r02_bench_config3_D6_chi144.json 2>/dev/null
timeout 600 python bench.py --D 7 --chi 196 --d 4 --nx 4 --ny 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_config5_D7_chi196_d4_n1.json 2>/dev/null
timeout 600 python bench.py --D 8 --chi 256 --nx 2 --ny 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_2x2.json 2>/dev/null
for f in gpurun_out/r02_bench_config*.json gpurun_out/r02_bench_n1_2x2.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    g = d.get("gpu_torch_baseline", {})
    print(sys.argv[1].split("/")[-1], "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "gpu_torch", g.get("value"), "x", g.get("speedup_of_value"), "cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
# ---- launch list of the timed sweep of bench.py itself (2x2 cell so that one sweep fits the capture)
L=$(timeout 300 python bench.py --nx 2 --ny 2 --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read())['gpu_launches'])")
echo "library launches per 2x2 sweep: $L"
SKIP=$((3 * L + 200))
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP -c 3200 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --nx 2 --ny 2 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_bench.csv")) if len(r) > 10 and r[0].isdigit()]
agg, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("ab200::", "").replace("(anonymous namespace)::", "")
    try:
        t = float(r[-1].replace(",", ""))
    except ValueError:
        continue
    unit = r[-2]
    t_us = t * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1.0)
    agg[name] += t_us
    cnt[name] += 1
tot = sum(agg.values())
with open("gpurun_out/r02_launches_bench_summary.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: launches of the TIMED sweep of bench.py (2x2 cell, first 3200 launches = one left+right phase); serialised, cold: compare shares\n")
    f.write("kernel,launches,total_us,share\n")
    for n, t in sorted(agg.items(), key=lambda kv: -kv[1]):
        f.write(f"{n},{cnt[n]},{t:.1f},{t / tot:.4f}\n")
print(open("gpurun_out/r02_launches_bench_summary.csv").read()[:2500])
PY
gzip -f gpurun_out/r02_launches_bench.csv
# ---- ncu --set full of the kernels that changed this round + the dominant one
for k in encode_big row_exp cholqr_chunk cholqr_factor "dgemm_dmma_kernel<64, 64, 32, 32, (bool)1, (bool)0" double_layer_fused_d8; do
  tag=$(echo "$k" | tr -c 'a-zA-Z0-9_' '_' | cut -c1-24)
  prog=tools/prof_k7.py; case "$k" in cholqr*) prog=tools/prof_tsqr.py;; double_layer*) prog=tools/enc_probe.py;; esac
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -o gpurun_out/r02_ncu_full_$tag python $prog > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
