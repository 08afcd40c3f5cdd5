#!/bin/bash
# run quick_bench against each prebuilt libkqgpu.so variant under query-engines_b200/variants/
for d in query-engines_b200/variants/*; do
  cp $d/libkqgpu.so query-engines_b200/libkqgpu.so
  export KQ_JIT_CACHE=off
  echo "== $d"; timeout 120 python tools/quick_bench.py 100000000 $1
done
