#!/bin/bash
# Occupancy sweep of the time-fused neuron-backward kernel ON the GPU box: rebuild per variant, time one training step's kernels.
set -u
for OCC in 4 5 6; do
  touch event_flow_b200/csrc/lif_conv_bwd_tc.cu
  make -j8 EXTRA="-DEF_PWW_OCC=$OCC" > /dev/null 2>&1 || { echo "build failed $OCC"; continue; }
  echo -n "OCC=$OCC: "
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file /tmp/l.csv python tools/profile_step.py train > /dev/null 2>&1
  python tools/summarize_launches.py /tmp/l.csv | grep -E "launches,|pointwise" | tr '\n' ' '; echo
done
