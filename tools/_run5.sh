mkdir -p gpurun_out
env | grep -i nccl
timeout 600 python bench.py > gpurun_out/s3_bench_final.json 2> gpurun_out/s3_bench_final.err
cat gpurun_out/s3_bench_final.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 34992 -c 11664 --csv --log-file gpurun_out/r01_launches_bench_k7.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s3_ncu_bench.log 2>&1
tail -2 gpurun_out/s3_ncu_bench.log | cut -c1-300
wc -l gpurun_out/r01_launches_bench_k7.csv
gzip -f gpurun_out/r01_launches_bench_k7.csv
