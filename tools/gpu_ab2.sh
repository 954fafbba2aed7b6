#!/bin/bash
# A/B helper (round 2, second session): named test files with per-test durations, then short bench lines of the named models.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab2.sh <tag> "<pytest args>" resnet50 vit_b16 ...'
tag=$1; shift
targs=$1; shift
timeout 600 python -m pytest $targs -m gpu -x -q --durations=12 > gpurun_out/tests_${tag}.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/tests_${tag}.log
for m in "$@"; do timeout 200 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_$m.json 2> gpurun_out/bench_${tag}_$m.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${tag}_$m.json'));print('$m',round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'], [(k['kernel'],k['ms']) for k in d.get('kernels',[])[:6]])"; done
