#!/bin/bash
# One gpurun call that validates and times the opt-in kernel variants written at the end of round 1
# (DESIGN.md 3.1 / 9).  Every step runs under its own `timeout`: the hybrid stem is new tcgen05 code (UMMA N = 32
# descriptors) that has never run, and a wrong descriptor shows up as an mbarrier wait that never returns -- it
# runs LAST so that a hang cannot cost the log-mel results.
#   gpurun --timeout 300 -- 'bash tools/try_experimental_variants.sh'
mkdir -p gpurun_out
out=gpurun_out/experimental_variants.txt
: > $out
echo "== log-mel packed variants: parity" >> $out
AFS_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_logmel.py -q -m gpu -k packed 2>&1 | tail -12 >> $out
for v in 0 1 2; do
  echo "== log-mel time, AFS_LOGMEL_PACKED=$v" >> $out
  AFS_LOGMEL_PACKED=$v timeout 60 python tools/time_logmel.py >> $out 2>&1
done
echo "== hybrid stem: parity (the stock stem / Conv64F / backbone tests with AFS_CONV1_HYBRID=1)" >> $out
AFS_CONV1_HYBRID=1 timeout 90 python -m pytest tests/test_gpu_models.py -q -m gpu -k "conv1 or conv64f or backbones" 2>&1 | tail -12 >> $out
for v in 0 1; do
  echo "== stem time, AFS_CONV1_HYBRID=$v" >> $out
  AFS_CONV1_HYBRID=$v timeout 45 python tools/run_conv1_tc.py >> $out 2>&1
done
cat $out
