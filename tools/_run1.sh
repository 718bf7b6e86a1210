set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s3_pytest.log 2>&1
tail -3 gpurun_out/s3_pytest.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
cat gpurun_out/s3_bench.json
timeout 300 python tools/trace_sweep.py --out gpurun_out/s3_trace.json > gpurun_out/s3_trace.txt 2>&1
head -40 gpurun_out/s3_trace.txt
for cfg in "0 0" "256 0" "192 0" "64 0" "128 3" "128 2"; do
  set -- $cfg
  ACETN_B200_I8_N1=$1 ACETN_B200_I8_STAGES=$2 timeout 200 python tools/i8_split_probe.py 2>&1 | tail -1 | tee -a gpurun_out/s3_split.txt
done
