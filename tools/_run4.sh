mkdir -p gpurun_out
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/s3_pytest4.log 2>&1
tail -3 gpurun_out/s3_pytest4.log
for n in 2; do timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/s3c_bench_n$n.json 2> gpurun_out/s3c_bench_n$n.err; python -c "
import json;d=json.loads(open('gpurun_out/s3c_bench_n$n.json').read().strip().splitlines()[-1]);print($n, d['value'], d['e2e'], d['config']['parallelism'])"; wc -l gpurun_out/s3c_bench_n$n.json; done
