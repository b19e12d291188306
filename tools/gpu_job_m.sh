#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/m_bench.json
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/m_ref.json
(timeout 300 python bench.py --envs 65536 --steps 10 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/m_bench65536.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_tick_quad -s 2301 -c 1 -f -o gpurun_out/r01_quad_v8_4096_odd python tools/prof_env.py 4096 2400 > gpurun_out/m_ncu1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 300 --csv --log-file gpurun_out/r01_launches_v8_4096.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_ncu4.log 2>&1
timeout 200 python tools/phase_tail.py 4096 2000 > gpurun_out/m_phase.log 2>&1
cut -c1-400 gpurun_out/m_bench.json; cut -c1-300 gpurun_out/m_ref.json; cut -c1-300 gpurun_out/m_bench65536.json; grep -E "^tick" gpurun_out/m_phase.log
