#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh r2u
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2)
