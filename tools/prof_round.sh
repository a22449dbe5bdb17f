#!/bin/bash
# Evidence for one round on the GPU box: tools/prof_round.sh <tag>   (outputs under gpurun_out/, summarised into profiles/ afterwards)
tag=${1:-r2w}
set -x
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
python bench.py --workload mha --steps 5 --warmup 3 > gpurun_out/${tag}_mha.json 2> gpurun_out/${tag}_mha.err
PLAIN="python bench.py --no-train --no-mha --no-cpu-baseline --no-eager-baseline --sustain-s 0 --steps 3 --warmup 3"
$PLAIN > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv $PLAIN > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"winattn_bwd|winattn_fwd|linbwd_tc_kernel|gemm_tc_kernel" --launch-skip 12 -c 6 -o gpurun_out/prof_${tag} -f $PLAIN > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log
