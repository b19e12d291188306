#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh r2q -k "other_cars or collision"
