#!/bin/bash
# Nsight Compute captures behind profiles/ (run under gpurun, ONE GPU):
#   bash tools/capture_profiles.sh <tag>
# 1. launch list of the bench command (one metric, every launch; cold-cache + serialised: compare SHARES)
# 2. `--set full` captures of the hot kernels from the short deterministic sequence tools/ncu_targets.py
# Summarise here (no GPU needed) with: python tools/summarize_ncu.py <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 5000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-ref-gpu > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
for spec in "gemm:gemm_kernel:9" "gemm:fa_fwd:2" "geom:render_:4" "geom:preprocess_kernel|run_sort|scatter_kernel|pack_kernel|lbs_skin|grid_:12" "geom:mlp_:8" "nn:gn_|layernorm:12"; do
    IFS=: read part pat cnt <<< "$spec"
    name=$(echo $pat | tr -c 'a-z0-9\n' '_' | cut -c1-24)
    timeout 300 $NCU --set full --import-source on -k "regex:$pat" -c $cnt -f -o $OUT/${TAG}_${name} \
        python tools/ncu_targets.py $part > $OUT/${TAG}_${name}.log 2>&1
    echo "$pat rc=$?"
done
# 3. DRAM traffic + duration of every tensor-core launch of ONE guidance pass (un-graphed, profiler-bracketed)
timeout 400 $NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    -k "regex:gemm_kernel|fa_fwd" --csv --log-file $OUT/${TAG}_gemm_traffic.csv python tools/gemm_pass.py > $OUT/${TAG}_gemm_pass.log 2>&1
echo "gemm traffic rc=$?"
# gpurun copies back at most 64 MiB: summarise HERE (ncu is on the box) and keep only the two reports whose source pages get read
mkdir -p $OUT/profiles_${TAG}
DWG_PROFILES_DIR=$OUT/profiles_${TAG} python tools/summarize_ncu.py $TAG
ls -la $OUT/*.ncu-rep
for f in $OUT/${TAG}_*.ncu-rep; do case "$f" in *gemm_kernel*|*render_*) ;; *) rm -f "$f" ;; esac; done
rm -f $OUT/${TAG}_gemm_traffic.csv
