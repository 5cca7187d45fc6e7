#!/bin/bash
# tools/ncu_capture.sh NAME KERNEL_REGEX SKIP -- command...   (run on the GPU box)
# One `ncu --set full` capture of the first launch matching KERNEL_REGEX after SKIP launches; the report is converted on the box
# into gpurun_out/NAME.raw.csv (+ NAME.source.csv) and deleted (reports with --import-source are ~40 MB, gpurun_out/ is capped at 64 MiB).
name=$1; regex=$2; skip=$3; shift 3; [ "$1" = "--" ] && shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -c 1 -s $skip -f -o /tmp/$name "$@" > gpurun_out/$name.ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>> gpurun_out/$name.ncu.log
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>> gpurun_out/$name.ncu.log
ls -la /tmp/$name.ncu-rep >> gpurun_out/$name.ncu.log
rm -f /tmp/$name.ncu-rep
