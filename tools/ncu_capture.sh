#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <bench args...>
# Captures one launch with ncu --set full on the GPU box and exports the raw
# and source pages as (gzipped) CSV into gpurun_out/ (the .ncu-rep itself is
# too large to travel back).
tag=$1; regex=$2; skip=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
    -o /tmp/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_${tag}_source.csv.gz
ls -la gpurun_out/ | grep $tag
