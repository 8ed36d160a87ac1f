#!/bin/bash
# ncu_export.sh <report.ncu-rep> : write <report>.raw.csv (+ <report>.source.csv.gz when the report has source)
# next to it and drop the report itself when it is larger than 20 MB (gpurun_out/ is capped at 64 MiB).
rep="$1"; base="${rep%.ncu-rep}"
ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
if [ "$2" == "source" ]; then ncu -i "$rep" --page source --csv 2>/dev/null | gzip -9 > "$base.source.csv.gz"; fi
sz=$(stat -c %s "$rep"); if [ "$sz" -gt 20000000 ]; then rm -f "$rep"; fi
