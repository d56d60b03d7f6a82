#!/bin/bash
# Evidence run for profiles/: launch list of a bench step, full ncu captures of the dominant kernels, compute-sanitizer logs.
# Usage (GPU box, from the repo root):  bash scripts/gpu_evidence.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# 1. launch list of one bench step at 20 000 pairs per length (per-launch times are cold-cache and serialised)
timeout 900 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $out/${tag}_launches_bench_20kpairs.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000 --no-e2e --no-cpu --no-swap --no-parity > $out/${tag}_launches_bench.log 2>&1
# 2. full captures: the cost-only kernel at the bench's 10 kb shape (25 000 pairs per launch) and the band fills of
#    interior-node-like pairs (4-state cells)
timeout 900 $NCU --set full --import-source on -k regex:k_cost_affine -c 1 -o $out/${tag}_k_cost_affine -f \
    python scripts/perf_probe.py --L 10000 --pairs 25000 --mode cost --reps 1 > $out/${tag}_ncu_cost.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_band2 -c 14 -o $out/${tag}_k_band2_4state -f \
    python scripts/perf_probe.py --L 2000 --pairs 20000 --decorated 1.0 --mode align --reps 1 > $out/${tag}_ncu_band2.log 2>&1
for f in ${tag}_k_cost_affine ${tag}_k_band2_4state; do
    [ -f $out/$f.ncu-rep ] && ncu -i $out/$f.ncu-rep --page raw --csv > $out/$f.raw.csv 2>/dev/null
done
# 3. compute-sanitizer
SAN=compute-sanitizer
timeout 1200 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_newkk.py -m gpu -x -q \
    -k "edge_cases or related_pairs or probe_fills or low_latency or speculative or linear_edge or columnwise or cuda_matches_golden or generic_fallback" \
    > $out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $out/${tag}_sanitizer_memcheck.log
timeout 1500 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_newkk.py -m gpu -x -q \
    -k "edge_cases or low_latency or speculative or many_wide_pairs or cuda_matches_golden" \
    > $out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> $out/${tag}_sanitizer_racecheck.log
tail -3 $out/${tag}_sanitizer_memcheck.log $out/${tag}_sanitizer_racecheck.log
ls -la $out | grep $tag
