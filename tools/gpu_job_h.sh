#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 300 --csv --log-file gpurun_out/h_launches.csv python tools/prof_env.py 4096 2260 > gpurun_out/h_ncu.log 2>&1
tail -2 gpurun_out/h_ncu.log
