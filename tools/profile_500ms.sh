#!/bin/bash
# One `ncu --set full` capture of the 500 ms streaming kernel + the launch list of the bench command (a re-capture after
# a kernel change; tools/profile_round.sh does every mode).   gpurun --timeout 900 -- 'bash tools/profile_500ms.sh r02b'
tag=${1:-rXX}
out=gpurun_out
common="--steps 3 --warmup 1 --skip-cpu-baseline --e2e-steps 1 --sustain-s 0 --skip-other-modes --skip-parity --cohort-subjects 0"
B="python bench.py --steps 5 --warmup 3 --subjects 8 --skip-cpu-baseline --e2e-steps 1 --sustain-s 0 --cohort-subjects 0"
$B > $out/plain_launches_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_launches_$tag.log 2>&1
C="python bench.py --mode 500ms --subjects 4 $common"
$C > $out/plain_500ms_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_500ms $C > $out/ncu_${tag}_500ms.log 2>&1
tail -1 $out/ncu_${tag}_500ms.log
ls -la $out/prof_${tag}_500ms.ncu-rep
