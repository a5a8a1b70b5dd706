#!/bin/bash
# One `ncu --set full` capture per distinct kernel of a UNet pass (config A, C or E), summarised into gpurun_out/.
#   bash scripts/ncu_full.sh A r02
set -u
CFG=${1:-A}; TAG=${2:-r02}
OUT=gpurun_out
python scripts/ncu_capture.py --config $CFG --list > $OUT/${TAG}_oplist_$CFG.txt || exit 1
python scripts/ncu_capture.py --config $CFG > $OUT/${TAG}_ncu_plain_$CFG.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/${TAG}_launches_config${CFG}.csv python scripts/ncu_capture.py --config $CFG > /dev/null 2>&1
OPS=$(python scripts/ncu_pick.py $OUT/${TAG}_launches_config${CFG}.csv $OUT/${TAG}_oplist_$CFG.txt) || exit 1
echo "config $CFG: full captures of ops $OPS"
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/${TAG}_full_$CFG \
    python scripts/ncu_capture.py --config $CFG --ops $OPS > $OUT/${TAG}_ncu_full_$CFG.log 2>&1
ncu -i /tmp/${TAG}_full_$CFG.ncu-rep --page raw --csv > /tmp/${TAG}_full_$CFG.csv
python scripts/ncu_summarize.py /tmp/${TAG}_full_$CFG.csv "ncu --set full --clock-control none, one launch per distinct kernel of a config-$CFG UNet pass (ops $OPS of ${TAG}_oplist_$CFG.txt; cold-cache, serialised)" > $OUT/${TAG}_ncu_full_config${CFG}.txt
ls -la /tmp/${TAG}_full_$CFG.ncu-rep
