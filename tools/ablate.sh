#!/bin/bash
# conv_tc ablations via RGFM_TC_DBG (1 = no epilogue work, 2 = no TMA loads, 4 = no b_full wait, 8 = no MMAs)
for d in 0 1 2 8 3 9 10 11; do
  echo "=== DBG=$d"
  RGFM_TC_DBG=$d RGFM_PROFILE_VERBOSE=1 timeout 120 python tools/profile_step.py --precision fp16 > gpurun_out/abl_$d.log 2>&1
  grep -E "^conv_tc|^total" gpurun_out/abl_$d.log
  grep -E "encoder_blocks.0.conv1 32x32 32->32|encoder_blocks.0.conv1 32x32 64->64|out_conv 32x32 32->1|encoder_blocks.3.conv1 16x16 128->128|middle_block1.conv1 8x8|input_conv 32x32 1->32" gpurun_out/abl_$d.log | sort -u -k5 | head -8
done
