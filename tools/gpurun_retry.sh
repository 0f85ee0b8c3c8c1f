#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (rc 3: no slot, nothing charged).
#   tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
# the library must load here before a box is spent on it (an undefined symbol costs a whole call)
python -c "import ctypes; ctypes.CDLL('$(dirname "$0")/../datashader_b200/libdsb200.so')" || { echo "libdsb200.so does not load"; exit 4; }
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout "$T" -- "$CMD" 2>&1); else out=$(gpurun --gpus "$G" --timeout "$T" -- "$CMD" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up after 40 transient answers"; exit 3
