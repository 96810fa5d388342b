# end-of-round-2 evidence run (1 GPU): full GPU suite, plain runs && ncu launch lists of the bench and of a training step,
# one --set full capture of the tall weight gradient, per-call op times, the parts of a step, the default bench line
set -o pipefail
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -8 > gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -4 > gpurun_out/r2g_smoke.log; tail -2 gpurun_out/r2g_smoke.log
BARGS="--steps 2 --warmup 3 --no-fullres --no-train --no-bandwidth --no-cudnn-baseline --no-cpu-baseline"
JPDSE_NO_GRAPH=1 python bench.py $BARGS > gpurun_out/r2g_plain_bench.log 2>&1 && JPDSE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py $BARGS > gpurun_out/r2g_ncu1.log 2>&1
python tools/one_train_step.py 1 > gpurun_out/r2g_plain_train.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/one_train_step.py 1 > gpurun_out/r2g_ncu2.log 2>&1
python tools/wgrad_probe.py conv3x3 2 32 64 1024 1024 5 > gpurun_out/r2g_plain_wgrad.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 2 -c 2 -o gpurun_out/r2_ncu_wgrad_tall_b2 python tools/wgrad_probe.py conv3x3 2 32 64 1024 1024 5 > gpurun_out/r2g_ncu3.log 2>&1
python tools/op_times.py --what d,vgg,g > gpurun_out/r2_op_times_d_vgg_g_b2.txt 2>&1
python tools/train_parts.py > gpurun_out/r2_train_parts_b2.txt 2>&1; tail -11 gpurun_out/r2_train_parts_b2.txt
bash tools/wgrad_shapes.sh > gpurun_out/r2_wgrad_shapes_b2.txt 2>&1
python tools/one_train_step.py 30 2>&1 | tail -1
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2g_bench.err; tail -c 600 gpurun_out/r2_final_bench.json
