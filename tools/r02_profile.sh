#!/bin/bash
# Round-2 GPU evidence, one bounded gpurun call (run from the repo root on one B200):
#   1. pytest -m gpu, smoke(), the default bench line (train + decode + decode512 + cpu_baseline)
#   2. ncu launch list of ONE training step of the same command
#   3. ncu --set full of the dominant kernels on single-layer probes at the benchmarked shapes (batch 64): a full report of
#      a whole training step is too slow to capture (ncu saves the step's ~100 GB of device memory per replayed kernel).
# Every stage has its own timeout; reports are summarised on the box and deleted (only small text files travel back).
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
O=gpurun_out
stamp pytest; timeout 400 python -m pytest tests -m gpu -x -q > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02h_pytest.log; tail -3 $O/r02h_pytest.log
stamp smoke; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02h_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02h_smoke.log
stamp bench; timeout 500 python bench.py > $O/r02h_bench.json 2> $O/r02h_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r02h_bench.json
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --no-parity"
stamp "launch list (train)"
VFM_CUDA_PROFILER_RANGE=train timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train.csv $B > $O/ncu_launch_train.log 2>&1; echo "rc=$?"
python tools/ncu_summarize.py launches $O/r02_launches_train.csv $O/r02_launches_train.md "ncu launch list of ONE f16d32 D-legacy training step (batch 64, 256x256, 1 GPU): $B, profiler range = the timed step"; gzip -f $O/r02_launches_train.csv; head -12 $O/r02_launches_train.md
K='regex:conv_tc_kernel|wgrad_tc_kernel|nhwc_prepass|upfirdn2d_blur|act_grad|gsum'
probe() {  # name, conv_probe arguments
  name=$1; shift
  stamp "set full: $name"
  timeout 170 ncu --set full --clock-control none --kernel-name-base demangled -k "$K" -c 10 -f -o $O/full_$name python tools/conv_probe.py --warm 0 --iters 1 "$@" > $O/ncu_full_$name.log 2>&1; echo "rc=$?"
  python tools/ncu_summarize.py full $O/full_$name.ncu-rep $O/r02_ncu_full_$name.md $O/r02_traffic_$name.json; rm -f $O/full_$name.ncu-rep
}
probe bwd_128_256 --cin 128 --cout 128 --res 256 --mode bwd
probe bwd_256_128 --cin 256 --cout 256 --res 128 --mode bwd
probe bwd_512_64 --cin 512 --cout 512 --res 64 --mode bwd
probe up2_256_128 --cin 256 --cout 128 --res 128 --up 2 --mode fused
rm -f $O/*.ncu-rep
stamp done; ls -la $O | head -40
