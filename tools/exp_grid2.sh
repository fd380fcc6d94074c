#!/bin/bash
for cfg in "4 3" "6 3" "8 3" "4 4" "4 6" "6 6" "8 6" "12 9" "16 12"; do set -- $cfg
  echo -n "C2 water_blocks=$1 solute_blocks=$2: "
  QNB_WATER_BLOCKS=$1 QNB_SOLUTE_BLOCKS=$2 timeout 200 python tools/exp_step.py C2 | sed 's/graph=True one_stream=0//'
done
for wb in 4 6 8 12; do echo -n "C5 water_blocks=$wb: "; QNB_WATER_BLOCKS=$wb timeout 200 python tools/exp_step.py C5 | sed 's/graph=True one_stream=0//'; done
