#!/bin/bash
# One command per piece of ncu evidence (run on the GPU box: `gpurun -- 'bash tools/ncu_stage.sh <what> [tag]'`).
# The library wraps the enqueue of every pipeline stage in an NVTX range (uvip_pyramid, uvip_fast, uvip_quadtree, uvip_blur,
# uvip_select, uvip_describe — nested in uvip_extract_group — and uvip_knn2, uvip_search_window), so a stage is selected by name, not by
# counting launches.
#   launches        launch list of one bench step at batch 256 (one stream): duration, DRAM bytes, warp instructions, pipes
#                   -> gpurun_out/<tag>_launches.csv, then regenerates profiles/kernel_pipes.json + roofline_traffic.json
#   full <stage>    ncu --set full --import-source on of the kernels inside NVTX range uvip/<stage> -> gpurun_out/<tag>_<stage>.ncu-rep
# Numbers printed by bench.py under ncu are never bench values.
set -e
what=${1:-launches}; shift || true
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
BENCH="python bench.py --batch 256 --steps 1 --warmup 3 --no-extras"
case "$what" in
  launches)
    tag=${1:-r2}
    # the last complete step of the run: skip the warm-up steps' launches (3 warm-up + 1 timed step, 14 launches each: import, 7 resizes,
    # fast, quadtree, blur, select, describe, knn2), keep the timed step and what follows
    UVIP_SERIAL=1 ncu --metrics $METRICS --clock-control none -s 40 -c 45 --csv --log-file gpurun_out/${tag}_launches.csv $BENCH > gpurun_out/${tag}_launches.log 2>&1
    python tools/pipes_from_launches.py gpurun_out/${tag}_launches.csv
    ;;
  full)
    stage=${1:?stage name: pyramid fast quadtree blur select describe knn2}; tag=${2:-r2}
    case "$stage" in knn2|search_window) expr="uvip_${stage}/";; *) expr="uvip_extract_group/uvip_${stage}/";; esac
    # the stage's kernels of ONE step (the second: the first is cold); reports of eight launches are 15-25 MB each and gpurun brings
    # back 64 MiB at most
    case "$stage" in pyramid) n=8;; *) n=1;; esac
    UVIP_SERIAL=1 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "$expr" -s $n -c $n \
        -o gpurun_out/${tag}_${stage} -f $BENCH > gpurun_out/${tag}_${stage}.log 2>&1
    python tools/ncu_summary.py gpurun_out/${tag}_${stage}.ncu-rep
    ;;
  *) echo "usage: $0 launches [tag] | full <stage> [tag]"; exit 2;;
esac
