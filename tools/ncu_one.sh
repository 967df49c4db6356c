#!/bin/bash
# One kernel launch under ncu --set full with the whole source page exported.
# usage: ncu_one.sh NAME REGEX SKIP -- cmd...
set -u
name=$1; regex=$2; skip=$3; shift 4
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name rc=$?"
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
