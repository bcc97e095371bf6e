for sb in 0 32 64 128; do
UVIP_SUBBATCH=$sb timeout 300 python bench.py --steps 50 --warmup 5 --no-extras > gpurun_out/r2i_sub$sb.json 2> gpurun_out/r2i_sub$sb.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_sub$sb.json").read().strip().splitlines()[-1])
print("subbatch $sb: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["stage_ms_per_step"])
PY
done
for sb in 0 32 64; do
UVIP_SERIAL=1 UVIP_SUBBATCH=$sb ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2i_traffic_sub$sb.csv python bench.py --batch 256 --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
done
