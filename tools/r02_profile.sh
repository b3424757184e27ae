#!/bin/bash
# Round-2 GPU evidence, one bounded gpurun call (run from the repo root on one B200):
#   1. pytest -m gpu, smoke(), the default bench line (train + decode + decode512 + cpu_baseline), a per-layer-shape kernel table
#   2. ncu --set full of the dominant kernels on single-layer probes at the benchmarked shapes (batch 64): a full report of a whole
#      training step is too slow to capture (tens of GB resident: ncu needs ~0.2 s per profiled launch even for one metric)
#   3. the reference's stock GPU path as context (tools/ref_gpu_bench.py)
#   4. ncu launch list of ONE training step of the bench command (last: it is the long one, ~3500 launches)
# Every stage has its own timeout; reports are summarised on the box and deleted (only small text files travel back).
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
O=gpurun_out; TAG=${1:-r02z}
stamp pytest; timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
stamp smoke; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
stamp bench; timeout 500 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json
stamp "bench, per-shape kernel table"; VFM_TIMING_DETAIL=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --no-parity > $O/${TAG}_bench_detail.json 2> $O/${TAG}_bench_detail.err; echo "rc=$?"
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
probe up2_640_32 --cin 640 --cout 512 --res 32 --up 2 --mode fused
# weights = layers of the benchmarked step a probe stands for (4 stride-1 layers per shape; 3 up=2 layers, two of them probed)
python tools/ncu_summarize.py merge $O/r02_traffic_train.json $O/r02_traffic_bwd_128_256.json:4 $O/r02_traffic_bwd_256_128.json:4 $O/r02_traffic_bwd_512_64.json:4 $O/r02_traffic_up2_256_128.json:1.5 $O/r02_traffic_up2_640_32.json:1.5
rm -f $O/*.ncu-rep
stamp "reference, stock GPU path (context)"; timeout 420 python tools/ref_gpu_bench.py --steps 3 --warmup 2 --json $O/r02_reference_stock_gpu.json > $O/ref_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/ref_gpu.log | cut -c1-400
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --no-parity"
stamp "launch list (train)"
VFM_CUDA_PROFILER_RANGE=train timeout ${LAUNCH_LIST_TIMEOUT:-900} ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train.csv $B > $O/ncu_launch_train.log 2>&1; echo "rc=$?"
python tools/ncu_summarize.py launches $O/r02_launches_train.csv $O/r02_launches_train.md "ncu launch list of ONE f16d32 D-legacy training step (batch 64, 256x256, 1 GPU): $B, profiler range = the timed step"; gzip -f $O/r02_launches_train.csv; head -14 $O/r02_launches_train.md
stamp done; ls -la $O | head -50
