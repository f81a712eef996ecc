#!/bin/bash
# usage: tools/ncu_export.sh <tag> <ncu -k/-s/-c args...> -- runs tools/profile_targets.py under ncu --set full and
# brings back only CSV exports (raw + source pages); the .ncu-rep stays on the box (gpurun_out is capped at 64 MiB)
tag=$1; shift
ncu --set full --clock-control none --import-source on "$@" -o /tmp/$tag python tools/profile_targets.py > gpurun_out/${tag}.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
ls -la /tmp/$tag.ncu-rep gpurun_out/${tag}_raw.csv gpurun_out/${tag}_src.csv
