#!/bin/bash
# compute-sanitizer over the kernels touched in the second half of round 2: K8 (shared-memory transposed atomics),
# K2 / K12 / K4d (programmatic dependent launch, coherent loads), K4 (merged tail waves)
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL='(k8 or wgrad or k12 or k2 or rmsnorm or k4d or k4_attention_vs_oracle) and not tc2 and not tc1 and not mma and not graph and not static_cache'
for tool in memcheck racecheck; do
  out=gpurun_out/r2b_sanitizer_${tool}.log
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 40 \
    python -m pytest tests/test_kernels_gpu.py -q -k "$SEL" -p no:cacheprovider > $out 2>&1
  echo "$tool rc=$?" | tee -a $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" $out | tail -6
done
