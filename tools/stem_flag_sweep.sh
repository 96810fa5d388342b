# role-switch sweep of rowconv_kernel<5,false> (the 7x7 stem) at the bench shape (JPDSE_DEBUG_FLAGS, see conv_rowstat.cu)
FLAGS=${FLAGS:-"0 1 2 3 8 16 24"}
for f in $FLAGS; do JPDSE_DEBUG_FLAGS=$f python tools/conv_probe.py stem 16 512 1024 40 64 10 2>&1 | tail -1; done
