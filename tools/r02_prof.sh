# usage: bash tools/r02_prof.sh TAG KERNEL_REGEX [workload] [count]   -- launch list + one full capture of the named kernel
TAG=$1; K=$2; W=${3:-c3}; C=${4:-1}
O=gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_$W.csv python tools/profile_step.py $W > $O/${TAG}_launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$K" -s 0 -c $C -f -o $O/${TAG}_$W python tools/profile_step.py $W > $O/${TAG}_full_$W.log 2>&1
tail -3 $O/${TAG}_full_$W.log
