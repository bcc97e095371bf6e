# end-to-end schedule sweep (bench.py e2e leg): handle pairs x batches in flight x chunk size
for cfg in "1 2 128" "2 2 128" "2 4 128" "2 4 64" "3 6 128" "2 4 256"; do
set -- $cfg
UVIP_E2E_HANDLES=$1 UVIP_E2E_INFLIGHT=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-extras --e2e-chunk $3 > gpurun_out/e2e_$1_$2_$3.json 2> gpurun_out/e2e_$1_$2_$3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_$1_$2_$3.json").read().strip().splitlines()[-1])
print("handles $1 inflight $2 chunk $3: e2e", round(d["e2e"]["value"]), "device", round(d["value"]), "link", round(d["e2e"]["link_ceiling_frames_per_s"]), d["e2e"]["chained_knn_equals_two_call_path"])
PY
done
UVIP_SERIAL=1 UVIP_SUBBATCH=32 timeout 300 python bench.py --steps 40 --warmup 5 --no-extras > gpurun_out/serial_sub32.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/serial_sub32.json').read().strip().splitlines()[-1]); print('serial subbatch 32 value', round(d['value']))"
UVIP_SERIAL=1 UVIP_SUBBATCH=64 timeout 300 python bench.py --steps 40 --warmup 5 --no-extras > gpurun_out/serial_sub64.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/serial_sub64.json').read().strip().splitlines()[-1]); print('serial subbatch 64 value', round(d['value']))"
UVIP_SERIAL=1 UVIP_SUBBATCH=64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/r2j_traffic_sub64.csv python bench.py --batch 256 --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
UVIP_SERIAL=1 UVIP_SUBBATCH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 28 -c 45 --csv --log-file gpurun_out/r2j_traffic_sub0.csv python bench.py --batch 256 --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
