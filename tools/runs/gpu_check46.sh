#!/bin/bash
# memcheck over the new kernels (incl. the long-context attention), then the round-end flow
bash tools/runs/gpu_check42.sh ${1:-r2fin4}_san
bash tools/gpu_final.sh ${1:-r2fin4}
