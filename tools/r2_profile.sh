#!/bin/bash
# Round-2 ncu captures (run under gpurun): every ncu command follows a plain run of the same command line.
P=gpurun_out
mkdir -p $P
NCU="ncu --clock-control none --profile-from-start off"
for w in "$@"; do
  case $w in
    c2|c4)
      python tools/prof_traversal.py $w > $P/plain_$w.log 2>&1 &&
      $NCU --set full --import-source on -f -o $P/r2_full_$w python tools/prof_traversal.py $w > $P/ncu_$w.log 2>&1 ;;
    c5)
      python tools/prof_traversal.py c5 > $P/plain_c5.log 2>&1 &&
      $NCU --set full --import-source on -k regex:k_trace_spec2_2l -s 1 -c 4 -f -o $P/r2_full_c5 python tools/prof_traversal.py c5 > $P/ncu_c5.log 2>&1
      python tools/prof_traversal.py c5 > $P/plain_c5.log 2>&1 &&
      $NCU --metrics gpu__time_duration.sum --csv --log-file $P/r2_launches_render_c5_8spp.csv python tools/prof_traversal.py c5 > $P/ncu_c5l.log 2>&1 ;;
    c3)
      python tools/prof_traversal.py c3 > $P/plain_c3.log 2>&1 &&
      $NCU --set full --import-source on -k regex:k_shade -s 5 -c 5 -f -o $P/r2_full_c3_shade python tools/prof_traversal.py c3 > $P/ncu_c3.log 2>&1
      python tools/prof_traversal.py c3 > $P/plain_c3.log 2>&1 &&
      $NCU --metrics gpu__time_duration.sum --csv --log-file $P/r2_launches_render_c3_16spp.csv python tools/prof_traversal.py c3 > $P/ncu_c3l.log 2>&1 ;;
    bench)
      python bench.py --steps 2 --warmup 1 > $P/plain_bench.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $P/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 > $P/ncu_bench.log 2>&1 ;;
  esac
done
# raw metric pages travel back as CSV; reports above ~10 MB stay on the box (gpurun_out/ is capped at 64 MiB)
for r in $P/*.ncu-rep; do
  [ -f "$r" ] || continue
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
  if [ $(stat -c %s $r) -gt 12000000 ]; then
    ncu -i $r --page source --csv > ${r%.ncu-rep}_source.csv 2>/dev/null
    gzip -f ${r%.ncu-rep}_source.csv
    rm -f $r
  fi
done
ls -la $P
