#!/bin/bash
# Runs on the GPU box: capture one kernel with ncu --set full, export the raw page and the
# hottest source lines as CSV (small), drop the .ncu-rep (large). usage: ncu_export.sh NAME REGEX SKIP COUNT -- cmd...
set -u
name=$1; regex=$2; skip=$3; count=$4; shift 5
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name rc=$?"
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/${name}_source_head.csv
ls -la /tmp/$name.ncu-rep
