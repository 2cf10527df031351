#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <gpus> -- retries a gpurun call of tools/_run.sh while the pod answers busy
T=${1:-1200}; G=${2:-1}
for i in $(seq 1 20); do
  if [ "$G" = 1 ]; then /usr/local/graft/bin/gpurun --timeout $T -- 'bash tools/_run.sh' > gpurun_out/_call.log 2>&1
  else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- 'bash tools/_run.sh' > gpurun_out/_call.log 2>&1; fi
  if grep -q "status=transient" gpurun_out/_call.log; then sleep 90; continue; fi
  break
done
grep -v "^+" gpurun_out/_call.log | tail -45
