#!/bin/bash
# The two ncu passes of /opt/skills/guides/B200_PROFILING.md over one bench step (run under gpurun, one GPU):
#   bash tools/run_profiles.sh r2c      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep
# then, in the authoring container:  python tools/summarize_ncu.py --tag <tag> --launches gpurun_out/<tag>_launches.csv \
#                                           --report gpurun_out/<tag>_prof.ncu-rep
TAG=${1:-prof}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:search_tc_kernel|resolve_records_kernel|step_kernel' -c 6 \
    -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_*; tail -2 gpurun_out/${TAG}_plain.log | cut -c1-300
