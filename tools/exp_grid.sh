#!/bin/bash
# sweep of blocks/SM of the two persistent kernels (design experiment)
for W in ${WORKLOADS:-C2 C3 C5}; do
for wb in 1 2 3 4; do for sb in 1 2 3; do
  echo -n "$W water_blocks=$wb solute_blocks=$sb: "
  QNB_WATER_BLOCKS=$wb QNB_SOLUTE_BLOCKS=$sb timeout 200 python tools/exp_step.py $W | sed 's/graph=True one_stream=0//'
  [ $W != C2 ] && break
done; done; done
