#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + `--set full` captures of the kernels of one training step,
# a 720p frame, an EDSR 1080p frame and the row-marching chain.  One gpurun call; plain runs first (exit 0 required).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r02}
python tools/profile_step.py 3 > gpurun_out/${tag}_profile_plain.log 2>&1 || { echo "plain profile workload failed"; tail -5 gpurun_out/${tag}_profile_plain.log; exit 1; }
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "launch list exit $?"
# skip the two warm-up repetitions' launches: capture the third repetition of every kernel
ncu --set full --clock-control none --import-source on \
    -k regex:'conv3x3_chain_kernel|conv3x3_row_kernel|wgrad_tc_kernel|wgrad_reduce_kernel|head_bicubic_kernel|head_wgrad|adamw_pack_kernel|conv3x3_tc_kernel' \
    --launch-skip 0 -c 48 -o gpurun_out/${tag}_step python tools/profile_step.py 1 > gpurun_out/${tag}_ncu_step.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/${tag}_step.ncu-rep
