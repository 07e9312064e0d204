#!/bin/bash
# One GPU-box session: parity tests, full-scale bench (both arms), ncu launch list of one bench command and
# one `--set full` capture of the top kernels.  Usage: tools/profile_round.sh <tag> [scale]
# Everything lands under gpurun_out/; summaries are written to profiles/ by tools/summarise_profiles.py here.
set -u
TAG=${1:-r01c}
SCALE=${2:-1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py --scale $SCALE > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"
cat $OUT/bench_$TAG.json
# launch list (cold-cache, serialised; shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --scale $SCALE --steps 1 --no-cpu-baseline --no-stress --no-decode > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
# full capture of the streaming kernels + the top latency kernels (first launch of each in the 4th step)
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'^(k1_classify|k1_compact|sd_block_stats|sd_resolve|k2_emit_pairs|max_span_kernel|ahc_replay|ahc_components)$' \
  --launch-count 8 -o $OUT/prof_top_$TAG -f python bench.py --scale $SCALE --steps 1 --no-cpu-baseline --no-stress --no-decode > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -8
