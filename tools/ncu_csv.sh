#!/bin/bash
# usage: tools/ncu_csv.sh <report.ncu-rep>   ->  <report>.raw.csv (one row per launch, all metrics) and removes the report
# (gpurun brings back at most 64 MiB: full reports with imported source do not fit, their CSV pages do)
set -e
rep="$1"
base="${rep%.ncu-rep}"
ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
if [ "$2" = "source" ]; then ncu -i "$rep" --page source --csv > "$base.source.csv" 2>/dev/null || true; fi
rm -f "$rep"
