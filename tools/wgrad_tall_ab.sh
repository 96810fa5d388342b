#!/bin/bash
# A/B of the tall (M = 256) weight-gradient items against the M = 128 form on the generator's shapes at batch 2 / 4
for shape in "conv3x3 2 32 64 1024 1024" "conv3x3 4 32 64 1024 1024" "convs2 2 64 128 512 1024" "convs2 2 128 256 256 512" \
             "convs2 2 256 512 128 256" "convt 2 32 64 1024 512" "convt 2 64 128 512 256" "convt 2 128 256 256 128"; do
  for t in 1 0; do
    echo -n "TALL=$t  "; JPDSE_WGRAD_TALL=$t python tools/wgrad_probe.py $shape 50 2>&1 | tail -1
  done
done
