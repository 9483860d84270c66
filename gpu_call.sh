mkdir -p gpurun_out
CMD="python scripts/prof_one.py 1 16 8192 64 0"
$CMD > gpurun_out/c10_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 8 -c 4 -f -o gpurun_out/c10_prof_d64 $CMD > gpurun_out/c10_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/c10_ncu.log
