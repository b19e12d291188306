#!/bin/bash
mkdir -p gpurun_out
L=projectd_core_b200
echo "baseline 65536 k_tick" > gpurun_out/r2k_q.log; tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "k_tick 96 regs" >> gpurun_out/r2k_q.log; PD_B200_LIB=$PWD/$L/libpd_b200_s10.so tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "quad8 168 regs, collision warp" >> gpurun_out/r2k_q.log; PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "quad8 168 regs, k_collide" >> gpurun_out/r2k_q.log; PD_COLL_WARP=0 PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "quad8 128 regs, collision warp" >> gpurun_out/r2k_q.log; PD_B200_LIB=$PWD/$L/libpd_b200_q5.so PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "quad8 128 regs, k_collide" >> gpurun_out/r2k_q.log; PD_B200_LIB=$PWD/$L/libpd_b200_q5.so PD_COLL_WARP=0 PD_QUAD_MAX_ENVS=100000 PD_QUAD_CPW=8 tools/quick_bench.sh 65536 >> gpurun_out/r2k_q.log 2>&1
echo "quad4 128 regs 4096" >> gpurun_out/r2k_q.log; PD_B200_LIB=$PWD/$L/libpd_b200_q5.so tools/quick_bench.sh 4096 >> gpurun_out/r2k_q.log 2>&1
echo "quad8 128 regs 8192" >> gpurun_out/r2k_q.log; PD_B200_LIB=$PWD/$L/libpd_b200_q5.so tools/quick_bench.sh 8192 >> gpurun_out/r2k_q.log 2>&1
echo "k_tick 32768 / 16384" >> gpurun_out/r2k_q.log; tools/quick_bench.sh 32768 >> gpurun_out/r2k_q.log 2>&1; tools/quick_bench.sh 16384 >> gpurun_out/r2k_q.log 2>&1
cat gpurun_out/r2k_q.log
