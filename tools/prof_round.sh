set -x
python bench.py --workload cfg5-sweep --steps 5 --warmup 3 > gpurun_out/r2r_sweep.json 2> gpurun_out/r2r_sweep.err
python bench.py --workload mha --steps 5 --warmup 3 > gpurun_out/r2r_mha.json 2> gpurun_out/r2r_mha.err
python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r2r_cfg5.json 2> gpurun_out/r2r_cfg5.err
python tools/profile_step.py cfg3 > gpurun_out/r2r_cfg3_kernels.md 2> gpurun_out/r2r_prof.err
python bench.py --no-train --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2r_plain.json 2> gpurun_out/r2r_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2r.csv python bench.py --no-train --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2r_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"winattn_bwd|winattn_fwd|linbwd_tc_kernel|gemm_tc_kernel" --launch-skip 12 -c 6 -o gpurun_out/prof_r2r -f python bench.py --no-train --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r2r_ncu2.log 2>&1
tail -2 gpurun_out/r2r_ncu2.log
