set -x
for c in k1 k1_125k k1_rk23 k2 k3 k4; do
  python tools/prof_one.py $c > gpurun_out/prof_one_$c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbk_run -s 1 -c 1 -f -o gpurun_out/prof_r02_$c python tools/prof_one.py $c > gpurun_out/ncu_r02_$c.log 2>&1
  tail -1 gpurun_out/prof_one_$c.log | cut -c1-200
done
