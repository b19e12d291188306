#!/bin/bash
# compute-sanitizer (memcheck) over the kernels added in round 2: k_collide2, the double-wishbone quad / serial instances, the device BVH build, contacts
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/san_small.log 2>&1; echo "sanitize_small rc=$?" > gpurun_out/san_summary.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "device_bvh or collision_flag or (multilink and quad) or brake_disc" > gpurun_out/san_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/san_summary.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_small.log gpurun_out/san_tests.log >> gpurun_out/san_summary.log
cat gpurun_out/san_summary.log
