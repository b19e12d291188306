#!/bin/bash
mkdir -p gpurun_out
PD_COLL_WARP=0 timeout 300 python tools/collide_tail.py > gpurun_out/r2v_collide.log 2>&1
tail -25 gpurun_out/r2v_collide.log
