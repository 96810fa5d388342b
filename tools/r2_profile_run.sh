set -o pipefail
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -8 > gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
python tools/op_times.py --what g > gpurun_out/r2f_op_times_g.txt 2>&1; grep -E "instnorm|====" gpurun_out/r2f_op_times_g.txt | head -24
BARGS="--steps 2 --warmup 3 --no-fullres --no-train --no-bandwidth --no-cudnn-baseline --no-cpu-baseline"
JPDSE_NO_GRAPH=1 python bench.py $BARGS > gpurun_out/r2f_plain_bench.log 2>&1 && JPDSE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py $BARGS > gpurun_out/r2f_ncu1.log 2>&1
python tools/one_train_step.py 1 > gpurun_out/r2f_plain_train.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/one_train_step.py 1 > gpurun_out/r2f_ncu2.log 2>&1
python tools/wgrad_probe.py conv3x3 2 32 64 1024 1024 5 > gpurun_out/r2f_plain_wgrad.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 2 -c 2 -o gpurun_out/r2_ncu_wgrad_b2 python tools/wgrad_probe.py conv3x3 2 32 64 1024 1024 5 > gpurun_out/r2f_ncu3.log 2>&1
python tools/fwd_bwd_time.py 2 3 > gpurun_out/r2f_plain_fb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:instnorm_backward_reduce -s 40 -c 2 -o gpurun_out/r2_ncu_reduce_b2 python tools/fwd_bwd_time.py 2 3 > gpurun_out/r2f_ncu4.log 2>&1
cat gpurun_out/r2f_plain_wgrad.log gpurun_out/r2f_plain_fb.log gpurun_out/r2f_plain_train.log | tail -5
python bench.py --steps 20 --warmup 3 --no-fullres --no-train --no-bandwidth --no-cudnn-baseline --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value', d['value'], 'e2e', d['e2e']['value'])"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches*.csv
