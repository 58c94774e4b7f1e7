#!/bin/bash
# One bench line per BASELINE.json configuration (run on a B200 box from the repository root):
#   bash profiles/run_workloads.sh <tag>        ->  gpurun_out/<tag>_bench_<workload>.json
# Copy the lines worth keeping to profiles/ (profiles/README.md lists them).
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_random.json 2> $out/${tag}_bench_random.err                                   # configs[2], the headline
python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err            # configs[1]
python bench.py --workload standing --steps 20 --warmup 5 --preroll 100 --no-cpu-baseline > $out/${tag}_bench_standing.json 2> $out/${tag}_bench_standing.err
python bench.py --workload rollout --no-cpu-baseline > $out/${tag}_bench_rollout.json 2> $out/${tag}_bench_rollout.err                   # configs[3]
: > $out/${tag}_bench_fallen_sweep.jsonl
for n in 1 2 3 4 5 6 7 8; do   # configs[4]; 64 timed steps = one whole episode (reset step included), starting on an episode boundary
  python bench.py --workload fallen --substeps $n --steps 64 --warmup 5 --preroll 59 --no-cpu-baseline >> $out/${tag}_bench_fallen_sweep.jsonl 2>> $out/${tag}_bench_fallen.err
done
python bench.py --no-contacts --steps 20 --warmup 5 --no-cpu-baseline --spread-steps 0 > $out/${tag}_bench_contact_free.json 2> $out/${tag}_bench_contact_free.err
