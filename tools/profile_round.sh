#!/bin/bash
# Round-end evidence, run on the GPU box:  bash tools/profile_round.sh <tag>
#   1. full GPU test suite, 2. bench line (own arm), 3. ncu launch list of the same bench command (cold-cache, serialised: shares only),
#   4. one ncu --set full capture of every kernel of a 128-frame step.  Numbers printed under ncu are never bench values.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
[ -z "$SKIP_TESTS" ] && python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/pytest_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err && tail -c 600 $out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --cpu-sample 0 > $out/ncu_launch_$tag.log 2>&1
# one ncu --set full capture per kernel of a 128-frame step (7 for the resize chain, 4 for the RANSAC waves), after the same
# command has exited 0 without ncu; only the raw-page CSV is kept (gpurun brings back at most 64 MiB)
python bench.py --frames 128 --steps 1 --warmup 1 --cpu-sample 0 > $out/plain_$tag.log 2>&1 || exit 1
: > $out/prof_${tag}_raw.csv
for spec in resize_tile:7 fast_strip:1 quadtree:1 blur_tile:1 describe:1 knn2:1 match_select:1 ransac_prepare:1 ransac_hyp_coop:2 ransac_hyp_kernel:3 ransac_select:1; do
  k=${spec%%:*}; n=${spec##*:}
  ncu --set full --clock-control none -k regex:$k -c $n -o $out/prof_${tag}_$k -f python bench.py --frames 128 --steps 1 --warmup 1 --cpu-sample 0 > $out/ncu_full_${tag}_$k.log 2>&1
  if [ -s $out/prof_${tag}_raw.csv ]; then ncu -i $out/prof_${tag}_$k.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 >> $out/prof_${tag}_raw.csv
  else ncu -i $out/prof_${tag}_$k.ncu-rep --page raw --csv 2>/dev/null > $out/prof_${tag}_raw.csv; fi
  rm -f $out/prof_${tag}_$k.ncu-rep
done
ls -la $out
