#!/bin/bash
# A/B of the single-level kernel's phase / refill thresholds: C2 microbench (closest / any-hit) and the C3 render.
P=gpurun_out
for t in 0 1 2 3; do
  echo "== B200PT_TRACE_TUNE=$t" | tee -a $P/r2_trace_tune.txt
  B200PT_TRACE_TUNE=$t python bench.py --no-cpu-baseline --no-e2e --no-c4-rays --path-iters 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('closest %.0f anyhit %.0f Mrays/s; C4 %.1f ms; C3 %.1f ms' % (d['closest_mrays'], d['anyhit_mrays'], d['path_tracing']['ms_per_image'], d['path_tracing_c3']['ms_per_image']))" | tee -a $P/r2_trace_tune.txt
done
