O=gpurun_out
echo "== correctness (VFM_ROW3=1)"
VFM_ROW3=1 timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_benchmark_config_gpu.py -m gpu -q -x -k "modulated_conv2d_vs_oracle or f16d32_layer_shapes_fp16 or fused_layer_matches or fused_training or fused_residual" 2>&1 | tail -6
echo "== timing"
for cfg in "--cin 128 --cout 128 --res 256" "--cin 256 --cout 256 --res 128" "--cin 512 --cout 512 --res 64"; do
  for m in fused bwd; do
    echo "-- $cfg $m"; timeout 60 python tools/conv_probe.py $cfg --mode $m --iters 5; VFM_ROW3=1 timeout 60 python tools/conv_probe.py $cfg --mode $m --iters 5
  done
done
