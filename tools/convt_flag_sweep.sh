# role-switch sweep of convt_fused_kernel on the two full-resolution up layers (JPDSE_DEBUG_FLAGS, see conv_convt.cu)
FLAGS=${FLAGS:-"0 1 2 3 8 16 24 32 48 56"}
for shape in "16 256 512 128 64" "16 128 256 256 128"; do
for f in $FLAGS; do JPDSE_DEBUG_FLAGS=$f python tools/conv_probe.py convt $shape 20 2>&1 | tail -1; done; done
