python -m pytest tests/test_traversal_gpu.py -x -q 2>&1 | tail -3
for v in 0 6 7 8; do echo "== variant $v"; python bench.py --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-path 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['closest_mrays'], d['anyhit_mrays'], d['roofline']['anyhit']['launch_ms'])"; done
