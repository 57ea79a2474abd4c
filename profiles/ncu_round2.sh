#!/bin/bash
# ncu evidence for round 2 (run on the GPU box through gpurun; one GPU).  Each capture is taken after the
# same command has run once without ncu.  Outputs: gpurun_out/r02_*_ncu_{raw.csv,details.txt} + launch lists.
set -u
mkdir -p gpurun_out
export CLIPB200_SYNTHETIC_WEIGHTS=1
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$rx" --launch-skip $skip --launch-count 1 \
       -f -o gpurun_out/${name} "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_ncu_raw.csv 2>/dev/null
  ncu -i gpurun_out/${name}.ncu-rep --page details > gpurun_out/${name}_ncu_details.txt 2>/dev/null
  echo "== $name: $(grep -c . gpurun_out/${name}_ncu_raw.csv) csv lines"
}
# 1. single-query search kernel over 10M rows
python profiles/search_probe.py 10000000 1 scan 1 > /dev/null 2>&1
cap r02_search_kernel "flatip_search_kernel<\\(int\\)1, \\(bool\\)1>" 2 python profiles/search_probe.py 10000000 1 scan 1
# 2. c_fc GEMM of the batch-256 vision tower (BN 256, bias+QuickGELU epilogue, CTA pairs)
python profiles/embed_launch_times.py 3 > /dev/null 2>&1
cap r02_gemm_cfc "gemm_tcgen05_kernel<\\(int\\)256, \\(int\\)1, \\(int\\)2>" 14 python profiles/embed_launch_times.py 3
# 3. tensor-core batch search, a 4M-row range of a batch of 1024 queries (ranges per search: 1K, 16K, 256K, 4M, 4M,
#    rest; two warm-up searches + one timed = 18 launches, the 16th is a 4M range)
python profiles/search_probe.py 10000000 1024 batch 1 > /dev/null 2>&1
cap r02_batch_kernel "flatip_batch_kernel<\\(int\\)2>" 15 python profiles/search_probe.py 10000000 1024 batch 1
# 3b. attention of the batch-256 vision tower (12 heads x 256 images, L = 50): the tcgen05 pair kernel, a launch
#     of a timed step (the first dozens belong to the calibration pass and the warm-up)
cap r02_attention_pair "attention_pair_kernel" 50 python profiles/embed_launch_times.py 3
# 4. launch lists (durations only)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_embed_launches.csv python profiles/embed_launch_times.py 3 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r02_search_launches.csv python profiles/search_probe.py 10000000 1,1024 auto 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
