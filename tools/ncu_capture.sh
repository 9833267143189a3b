#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> <bench args...>
# ncu --set full capture of one kernel family; keeps text exports (raw / details / source CSV) under
# gpurun_out/ and drops the .ncu-rep when it is too large to travel back (64 MiB cap on gpurun_out).
tag=$1; regex=$2; skip=$3; count=$4; shift 4
out=gpurun_out/prof_${tag}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o $out "$@" > gpurun_out/ncu_${tag}.log 2>&1
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
ncu -i $out.ncu-rep --page source --csv 2>/dev/null | gzip > $out.source.csv.gz
sz=$(stat -c %s $out.ncu-rep)
if [ "$sz" -gt 12000000 ]; then rm -f $out.ncu-rep; fi
ls -la gpurun_out/ | grep prof_${tag}
