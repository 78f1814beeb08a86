#!/bin/bash
# One gpurun call: GPU tests, launch list of the default bench command, one `ncu --set full` capture per analysis mode.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh v12'
tag=${1:-vX}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 5 --warmup 3 --subjects 8 --skip-cpu-baseline --e2e-steps 1 --clock-probe-s 0"
$B > $out/plain_launches_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_launches_$tag.log 2>&1
for m in 500ms 1s 2s; do
  C="python bench.py --mode $m --subjects 4 --steps 3 --warmup 1 --skip-cpu-baseline --e2e-steps 1 --clock-probe-s 0 --skip-other-modes --skip-parity"
  $C > $out/plain_$m.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:de_psd -s 3 -c 1 -f -o $out/prof_${tag}_$m $C > $out/ncu_${tag}_$m.log 2>&1
  tail -1 $out/ncu_${tag}_$m.log
done
