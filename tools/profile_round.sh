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
python bench.py --frames 128 --steps 1 --warmup 1 --cpu-sample 0 > $out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none -c 64 -o $out/prof_$tag -f python bench.py --frames 128 --steps 1 --warmup 1 --cpu-sample 0 > $out/ncu_full_$tag.log 2>&1
# gpurun brings back at most 64 MiB: keep the CSV of the raw page, drop the report
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null && rm -f $out/prof_$tag.ncu-rep
ls -la $out
