#!/bin/bash
# Turn a .ncu-rep into compact text (raw metrics csv + per-instruction source csv), gzip, and drop the report:
# gpurun only copies back 64 MiB.   usage: ncu_export.sh <report-without-extension>
R=$1
ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
ncu -i $R.ncu-rep --page source --csv > $R.source.csv 2>/dev/null
gzip -f $R.raw.csv $R.source.csv
rm -f $R.ncu-rep
ls -la $R.*
