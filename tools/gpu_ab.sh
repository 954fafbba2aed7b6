#!/bin/bash
# A/B helper used throughout round 1: full GPU test suite, then bench lines of the named models into gpurun_out/.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh <tag> resnet50 vit_b16 convnext_tiny swin_tiny'
tag=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in "$@"; do timeout 300 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_$m.json 2> gpurun_out/bench_${tag}_$m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${tag}_$m.json'));print('$m',round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'])"; done
