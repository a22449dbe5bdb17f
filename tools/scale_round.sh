#!/bin/bash
# weak-scaling runs of the default bench line (cfg2 kernel metric + nested cfg3 training step) and the cfg4 training step
# on one 8-GPU box: tools/scale_round.sh <tag>
tag=${1:-r2}
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --workload cfg4 --steps 20 --warmup 3 > gpurun_out/${tag}_cfg4_${n}gpu.json 2> gpurun_out/${tag}_cfg4_${n}gpu.err
done
python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/${tag}_cfg4_1gpu.json 2> gpurun_out/${tag}_cfg4_1gpu.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/TAG_*gpu.json".replace("TAG", "%s"))):
    pass
P
for f in gpurun_out/${tag}_bench_*gpu.json gpurun_out/${tag}_cfg4_*gpu.json; do python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); t=d.get('train') or {}
    print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'train', round(t.get('ms_per_step',0),2), round(t.get('samples_per_s',0),1), 'train_e2e', round((t.get('e2e') or {}).get('samples_per_s',0),1))
except Exception as e: print(sys.argv[1], 'ERR', e)
" $f; done
