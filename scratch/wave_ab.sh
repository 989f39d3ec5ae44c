for w in 25 26 27; do
  echo "== wave 2^$w"
  B200PT_WAVE_LOG2=$w python tools/run_config.py c3 --li 0 --crop 0 --reps 3 2>&1 | grep "^render\|rror"
done
echo "== c4 wave 26"
B200PT_WAVE_LOG2=26 python tools/run_config.py c4 --li 0 --crop 0 --reps 2 2>&1 | grep "^render\|rror"
echo "== c5 wave 22 / 25"
B200PT_WAVE_LOG2=22 python tools/run_config.py c5 --li 0 --crop 0 --reps 2 2>&1 | grep "^render\|rror\|preprocess"
B200PT_WAVE_LOG2=25 python tools/run_config.py c5 --li 0 --crop 0 --reps 2 2>&1 | grep "^render\|rror"
