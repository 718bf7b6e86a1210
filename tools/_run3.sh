mkdir -p gpurun_out
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s3_pytest3.log 2>&1
tail -3 gpurun_out/s3_pytest3.log
for w in 1 4 8; do
ACETN_B200_K2_WAVES=$w timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3_bench_w$w.json 2>> gpurun_out/s3_bench.err
python -c "import json;d=json.load(open('gpurun_out/s3_bench_w$w.json'));print('waves $w',d['value'],d['e2e']['value'])"
done
timeout 300 python tools/trace_sweep.py --out gpurun_out/s3_trace3.json > gpurun_out/s3_trace3.txt 2>&1
head -20 gpurun_out/s3_trace3.txt
