for k in 16 64; do
  cp pbrt-v3-rs_b200/libb200pt.so /tmp/lib_default.so
  cp scratch/libb200pt_k$k.so pbrt-v3-rs_b200/libb200pt.so
  echo "== kSmall $k"; python tools/bench_build.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['mesh'], 'gpu_ms', round(d['gpu_ms'],2), 'identical', d['identical'])"
  cp /tmp/lib_default.so pbrt-v3-rs_b200/libb200pt.so
done
echo "== kSmall 32 (default)"; python tools/bench_build.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['mesh'], 'gpu_ms', round(d['gpu_ms'],2), 'identical', d['identical'])"
