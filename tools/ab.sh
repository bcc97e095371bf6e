#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh <lib.so> [<lib.so> ...]  (variants built with build_cuda(extra=[-D...], out=...))
# prints the one-stream stage times and the two-stream value of each
for lib in "$@"; do
UVIP_LIB=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-extras > /tmp/ab.json 2>/tmp/ab.err || { echo "$lib FAILED"; tail -3 /tmp/ab.err; continue; }
python - "$lib" <<PY
import json, sys
d=json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1]); s=d["roofline"]["stage_ms_per_step"]
print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fast %.4f pyr %.4f qt %.4f blur %.4f desc %.4f knn %.4f" % (s["fast"], s["pyramid"], s["quadtree"], s["blur"], s["describe"], d["roofline"]["knn_ms_per_step"]))
PY
done
