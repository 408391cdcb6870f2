#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python tools/one_c4.py c4 > gpurun_out/r02_launches_c4.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_c4.csv 2>/dev/null | head -16
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_posterior.csv python tools/one_c4.py post > gpurun_out/r02_launches_post.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_posterior.csv 2>/dev/null | head -14
