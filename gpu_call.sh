mkdir -p gpurun_out; rm -f gpurun_out/c24_ab.jsonl
timeout 300 python scripts/dev_check.py --what all --cases small,mid,c2 > gpurun_out/c24_check.log 2>&1; echo "check rc=$?"; grep -E "EXCEPTION|hang" gpurun_out/c24_check.log | cut -c1-300; grep -c '"ok": true' gpurun_out/c24_check.log
for round in 1 2; do for v in old elect; do
FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so timeout 200 python scripts/ab_time.py all >> gpurun_out/c24_ab.jsonl 2>> gpurun_out/c24_ab.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/c24_ab.jsonl'):
    d=json.loads(l); print(d['lib'][11:-3], {k:(v['fwd'],v['dQ'],v['dKV']) for k,v in d.items() if isinstance(v,dict)})
PY
tail -3 gpurun_out/c24_ab.err
