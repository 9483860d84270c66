mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/c19_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c19_launches_c2.csv $CMD > gpurun_out/c19_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/c19_plain_c2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -f -o gpurun_out/c19_prof_c2 $CMD > gpurun_out/c19_ncu2.log 2>&1
echo "full c2 rc=$?"
CMD3="python bench.py --steps 2 --warmup 3 --no-extras --workload C3"
$CMD3 > gpurun_out/c19_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -f -o gpurun_out/c19_prof_c3 $CMD3 > gpurun_out/c19_ncu3.log 2>&1
echo "full c3 rc=$?"
