#!/bin/bash
# re-capture of the thread-per-car kernel at 65536 envs after a change to it (ncu --set full, summarised on the box)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:'^k_tick' -s 1300 -c 1 -f -o gpurun_out/r02g_serial_65536 python tools/prof_env.py 65536 1400 > gpurun_out/pf_r02g.log 2>&1
python tools/ncu_summary.py gpurun_out/r02g_serial_65536.ncu-rep "k_tick (thread per car), 65536 envs, round 2 final code with the thermal grids swept in place" > gpurun_out/r02g_serial_65536.md 2>> gpurun_out/pf_r02g.log
rm -f gpurun_out/r02g_serial_65536.ncu-rep
grep -i "dram__bytes\|local" gpurun_out/r02g_serial_65536.md | head -20
