#!/bin/bash
# bench.py under torchrun at N GPUs of one box + the C++ sequence example on the same GPUs.
#   gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_scale.sh N tag'
set -u
N=${1:-2}
TAG=${2:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 20 --warmup 5 \
    > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_${N}gpu_$TAG.err | cut -c1-200
python -c "
import numpy as np, sys
sys.path.insert(0,'.')
from deplex_b200 import synth
k=synth.intrinsics_for(480,640)
np.stack([synth.depth_to_cloud(synth.make_depth(480,640,i,k),k,'rowmajor') for i in range(64)]).tofile('/tmp/clouds.bin')
"
deplex_b200/cpp/build/process_sequence --clouds /tmp/clouds.bin 480 640 - 256 3 16 > $OUT/cpp_sequence_${N}gpu_$TAG.txt 2>&1; cat $OUT/cpp_sequence_${N}gpu_$TAG.txt
