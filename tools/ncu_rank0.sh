#!/bin/bash
# torchrun wrapper: rank 0 runs under ncu (duration-only launch list), the others plainly.
# Usage: python -m torch.distributed.run ... --no-python tools/ncu_rank0.sh <out.csv> script.py args...
out=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out" python "$@"
else
  exec python "$@"
fi
