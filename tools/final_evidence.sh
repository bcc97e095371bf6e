# end-of-round evidence on one B200: launch list (+ regenerated pipe / traffic JSON), full captures of the stages, sanitizer passes, bench lines
set -x
bash tools/ncu_stage.sh launches r2_final
cp profiles/kernel_pipes.json profiles/roofline_traffic.json gpurun_out/
for st in fast describe pyramid blur quadtree knn2; do bash tools/ncu_stage.sh full $st r2_final > gpurun_out/r2_final_${st}_full.csv 2>/dev/null; done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_final_memcheck.log 2>&1; tail -4 gpurun_out/r2_final_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg1_euroc_frame or batch_matches_single or chains_knn or hard_images" > gpurun_out/r2_final_racecheck.log 2>&1; tail -4 gpurun_out/r2_final_racecheck.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; tail -c 400 gpurun_out/r2_final_reference_arm.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -3 gpurun_out/r2_final_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("FINAL value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["e2e"]["frac_of_min_device_link"], "issue", d["roofline"].get("issue",{}).get("frac"), "traffic x", d["roofline"].get("traffic_all_stages",{}).get("vs_algorithmic"))
PY
