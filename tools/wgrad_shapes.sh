#!/bin/bash
# every weight-gradient shape of the generator at batch 2 (memset + wgrad + finalize per call)
for shape in "stem 2 512 1024 40 64" "head 2 512 1024 64 3" "convs2 2 512 1024 64 128" "convs2 2 256 512 128 256" "convs2 2 128 256 256 512" \
             "convs2 2 64 128 512 1024" "conv3x3 2 32 64 1024 1024" "convt 2 32 64 1024 512" "convt 2 64 128 512 256" \
             "convt 2 128 256 256 128" "convt 2 256 512 128 64"; do
  python tools/wgrad_probe.py $shape 30 2>&1 | tail -1
done
