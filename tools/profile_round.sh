#!/bin/bash
# One gpurun call: launch list of the default bench command + one `ncu --set full` capture per kernel of interest.
#   gpurun --timeout 1700 -- 'bash tools/profile_round.sh r02'
# Each ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
tag=${1:-rXX}
out=gpurun_out
common="--steps 3 --warmup 1 --skip-cpu-baseline --e2e-steps 1 --sustain-s 0 --skip-other-modes --skip-parity --cohort-subjects 0"
B="python bench.py --steps 5 --warmup 3 --subjects 8 --skip-cpu-baseline --e2e-steps 1 --sustain-s 0 --cohort-subjects 0"
$B > $out/plain_launches_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_launches_$tag.log 2>&1
for m in 500ms 1s 2s; do
  C="python bench.py --mode $m --subjects 4 $common"
  $C > $out/plain_$m.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_$m $C > $out/ncu_${tag}_$m.log 2>&1
  tail -1 $out/ncu_${tag}_$m.log
done
# 2 s mode with one TMA tensor copy per tile (measurement option): DRAM traffic against the bulk-copy loader
C="python bench.py --mode 2s --subjects 4 --tensor-loads $common"
$C > $out/plain_2s_tensor.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_2s_tensor $C > $out/ncu_${tag}_2s_tensor.log 2>&1
tail -1 $out/ncu_${tag}_2s_tensor.log
# pre-cut 500 ms windows (eegfe_de_psd_windows) and the GLMNet product (NORM instantiation)
C="python tools/win_bench.py --rows 2430400"
$C > $out/plain_win100.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_win100 $C > $out/ncu_${tag}_win100.log 2>&1
tail -1 $out/ncu_${tag}_win100.log
C="python tools/glmnet_bench.py --subjects 4"
$C > $out/plain_glmnet.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_glmnet $C > $out/ncu_${tag}_glmnet.log 2>&1
tail -1 $out/ncu_${tag}_glmnet.log
ls -la $out/*.ncu-rep
