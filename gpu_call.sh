for i in 1 2; do timeout 100 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('C2', round(d['value'],1), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done
