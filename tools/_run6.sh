mkdir -p gpurun_out
for cfg in "2 20 2" "4 64 2" "6 144 2" "7 196 4" "7 196 8" "8 128 2"; do
  set -- $cfg
  timeout 280 python bench.py --D $1 --chi $2 --d $3 --no-cpu-baseline > gpurun_out/s4_shape_$1_$2_$3.json 2> gpurun_out/s4_shape.err
  python -c "
import json;d=json.load(open('gpurun_out/s4_shape_$1_$2_$3.json'));print('D=$1 chi=$2 d=$3', round(d['value'],3), 'sweeps/s', round(d['roofline']['whole_sweep_tflops'],2),'TF', d['config']['thin_engine'][:4])" || tail -3 gpurun_out/s4_shape.err
done
