#!/bin/bash
# Round-2 profile capture, run on the GPU box:  gpurun -- 'bash tools/profile_r2.sh'
# Every ncu command is preceded by the same command without ncu (must exit 0); numbers printed
# under ncu are never bench values.
set -u
O=gpurun_out
B="python bench.py --pairs 131072 --steps 2 --warmup 1 --no-cpu --no-files --no-extras --no-capped"
$B > $O/r2p_bench_131k.json 2> $O/r2p_bench_131k.err || { echo "plain run failed"; exit 1; }
# (1) launch list of the same command
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2p_launches.csv $B \
  > $O/r2p_ncu_launches.log 2>&1
# (2) full capture of one chunk's kernels (second timed step: skip the first two passes' launches)
ncu --set full --clock-control none --import-source on --launch-skip 64 --launch-count 34 \
  -o $O/r2p_kernels $B > $O/r2p_ncu_kernels.log 2>&1
# (3) the screen kernels on the config-3 workload (80 % decoys)
W="python bench.py --workload wes_2x100 --pairs 1048576 --steps 1 --warmup 1 --no-cpu --no-files --no-extras --no-capped"
$W > $O/r2p_bench_wes_1M.json 2> $O/r2p_bench_wes_1M.err || { echo "plain wes run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_screen|k_prescreen" --launch-skip 16 -c 4 \
  -o $O/r2p_screen $W > $O/r2p_ncu_screen.log 2>&1
ls -la $O/r2p_*
