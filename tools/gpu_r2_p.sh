#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_car_parity.py ks_mazda_rx7_tuned 48 400 > gpurun_out/r2p_dbg.log 2>&1
tail -60 gpurun_out/r2p_dbg.log
