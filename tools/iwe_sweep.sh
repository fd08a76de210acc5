#!/bin/bash
# Tuning sweep of the event-warping loss kernels ON the GPU box: rebuild iwe_loss.cu per variant, time both windows.
set -u
for V in "-DEF_IWE_HINT=1" "-DEF_IWE_HINT=0" "-DEF_IWE_HINT=1 -DEF_IWE_EPT=2 -DEF_IWE_OCC_F=3 -DEF_IWE_OCC_B=3"; do
  touch event_flow_b200/csrc/iwe_loss.cu
  make -j8 EXTRA="$V" > /dev/null 2>&1 || { echo "build failed $V"; continue; }
  echo -n "$V  regs: "; grep -A2 "iwe_loss_fwd_kernel\|iwe_loss_bwd_kernel" build/iwe_loss.ptxas.log | grep -oE "Used [0-9]+ registers|[0-9]+ bytes spill stores" | tr '\n' ' '
  python tools/iwe_probe.py child 2>&1 | grep RESULT
done
