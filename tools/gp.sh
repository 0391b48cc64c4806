#!/bin/bash
# retry gpurun while the pod answers busy (exit 3 / status=transient); usage: tools/gp.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -${GP_TAIL:-25}; exit $rc
done
echo "gp.sh: gave up (pod busy)"; exit 3
