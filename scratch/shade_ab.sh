for b in 4 5 6; do
  echo "== shade blocks $b"
  B200PT_SHADE_BLOCKS=$b python tools/run_config.py c3 --li 0 --crop 0 --reps 3 2>&1 | grep "^render"
done
