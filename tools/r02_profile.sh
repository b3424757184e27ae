#!/bin/bash
# ncu evidence for round 2 (run under gpurun on one B200): launch list of one training step + `--set full` of the dominant kernels.
# Every capture is bounded (-c, timeout) and each report is summarised on the box and deleted (a full report is ~100 MB).
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
D="python bench.py --mode decode --steps 1 --warmup 3 --no-cpu-baseline --cuda-graph off"
NB="--kernel-name-base demangled"
$B > gpurun_out/plain_train.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
echo "== launch list (train)"
VFM_CUDA_PROFILER_RANGE=train timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv $B > gpurun_out/ncu_launch_train.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r02_launches_train.csv gpurun_out/r02_launches_train.md "ncu launch list of ONE f16d32 D-legacy training step (batch 64, 256x256, 1 GPU): bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline, profiler range = the timed step" && gzip -f gpurun_out/r02_launches_train.csv
echo "== set full: fp16 tensor-core kernels of the training step"
VFM_CUDA_PROFILER_RANGE=train timeout 500 ncu --profile-from-start off --set full --clock-control none $NB -k regex:"conv_tc_kernel<__half|wgrad_tc_kernel<false" -c 45 -f -o gpurun_out/full_tc $B > gpurun_out/ncu_full_train_tc.log 2>&1
python tools/ncu_summarize.py full gpurun_out/full_tc.ncu-rep gpurun_out/r02_ncu_full_train_tc.md gpurun_out/r02_traffic_train_tc.json; rm -f gpurun_out/full_tc.ncu-rep
echo "== set full: HBM-bound kernels of the training step"
VFM_CUDA_PROFILER_RANGE=train timeout 500 ncu --profile-from-start off --set full --clock-control none $NB -k regex:"nhwc_prepass_kernel<__half|act_grad_gsum|upfirdn2d_blur|rows_affine_kernel<__half|gn_bwd_reduce_kernel<__half|bias_act_rows_kernel<__half" -c 40 -f -o gpurun_out/full_hbm $B > gpurun_out/ncu_full_train_hbm.log 2>&1
python tools/ncu_summarize.py full gpurun_out/full_hbm.ncu-rep gpurun_out/r02_ncu_full_train_hbm.md gpurun_out/r02_traffic_train_hbm.json; rm -f gpurun_out/full_hbm.ncu-rep
echo "== set full: fp16 convs of the decode step"
$D > gpurun_out/plain_decode.log 2>&1 || { echo "plain decode run failed"; exit 1; }
VFM_CUDA_PROFILER_RANGE=decode timeout 400 ncu --profile-from-start off --set full --clock-control none $NB -k regex:"conv_tc_kernel<__half|upfirdn2d_blur" -c 24 -f -o gpurun_out/full_dec $D > gpurun_out/ncu_full_decode.log 2>&1
python tools/ncu_summarize.py full gpurun_out/full_dec.ncu-rep gpurun_out/r02_ncu_full_decode.md gpurun_out/r02_traffic_decode.json; rm -f gpurun_out/full_dec.ncu-rep
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | head -30
tail -2 gpurun_out/ncu_launch_train.log gpurun_out/ncu_full_train_tc.log gpurun_out/ncu_full_train_hbm.log gpurun_out/ncu_full_decode.log
