mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c29_bench.json')); print(d['value'], d['clocks'])"
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/c29_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c29_launches_c2.csv $CMD > gpurun_out/c29_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/c29_plain_c2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -f -o gpurun_out/c29_prof_c2 $CMD > gpurun_out/c29_ncu2.log 2>&1
echo "full c2 rc=$?"
CMD3="python bench.py --steps 2 --warmup 3 --no-extras --workload C3"
$CMD3 > gpurun_out/c29_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -f -o gpurun_out/c29_prof_c3 $CMD3 > gpurun_out/c29_ncu3.log 2>&1
echo "full c3 rc=$?"
