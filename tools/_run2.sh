set -x
mkdir -p gpurun_out
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s3_pytest2.log 2>&1
tail -3 gpurun_out/s3_pytest2.log
ACETN_B200_STAGGER=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3_bench_nostagger.json 2> gpurun_out/s3_bench.err
python -c "import json;d=json.load(open('gpurun_out/s3_bench_nostagger.json'));print('nostagger',d['value'],d['e2e']['value'])"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3_bench_stagger.json 2>> gpurun_out/s3_bench.err
python -c "import json;d=json.load(open('gpurun_out/s3_bench_stagger.json'));print('stagger',d['value'],d['e2e']['value'])"
timeout 300 python tools/trace_sweep.py --out gpurun_out/s3_trace2.json > gpurun_out/s3_trace2.txt 2>&1
head -20 gpurun_out/s3_trace2.txt
