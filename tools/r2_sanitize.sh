#!/bin/bash
# compute-sanitizer over the whole path (VERDICT r1 missing #7): memcheck, racecheck, synccheck, pair mode forced on / off.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for pair in 2 0; do
    out=gpurun_out/r2_sanitizer_${tool}_pair${pair}.log
    (RAMNET_PAIR=$pair timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_target.py 2>&1 | tail -60) > $out
    echo "== $tool pair=$pair: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' $out | tr '\n' ' ')"
  done
done
