#!/bin/bash
# final round-2 evidence at the final code: ncu --set full of the tick kernels and k_collide2 (summarised on the box: the reports are too large to
# travel back together), launch lists, phase clocks
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name kernel-regex skip envs ticks title
  timeout 500 $NCU -k regex:$2 -s $3 -c 1 -f -o gpurun_out/$1 python tools/prof_env.py $4 $5 > gpurun_out/pf_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep "$6" > gpurun_out/$1.md 2>> gpurun_out/pf_$1.log
  rm -f gpurun_out/$1.ncu-rep
}
cap r02f_quad_4096_odd k_tick_quad 2301 4096 2400 "k_tick_quad<4>, 4096 envs, odd frame (collision warp active), round 2 final code"
cap r02f_quad_4096_even k_tick_quad 2300 4096 2400 "k_tick_quad<4>, 4096 envs, even frame, round 2 final code"
cap r02f_serial_65536 '^k_tick' 1300 65536 1400 "k_tick (thread per car), 65536 envs, round 2 final code"
cap r02f_collide2_65536 k_collide2 650 65536 1400 "k_collide2<4> (floor test 4 lanes per car, warp for wall cars), 65536 envs, round 2 final code"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2050 -c 200 --csv --log-file gpurun_out/r02f_launches_4096.csv python tools/prof_env.py 4096 2300 > gpurun_out/pf_ncu3.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1950 -c 150 --csv --log-file gpurun_out/r02f_launches_65536.csv python tools/prof_env.py 65536 1400 > gpurun_out/pf_ncu4.log 2>&1
timeout 200 python tools/phase_tail.py 4096 2000 > gpurun_out/pf_phase.log 2>&1
timeout 300 python -m pytest tests/test_pyprojectd_dropin.py -q -m gpu > gpurun_out/pf_dropin.log 2>&1
tail -5 gpurun_out/pf_dropin.log; ls -la gpurun_out | tail -15
