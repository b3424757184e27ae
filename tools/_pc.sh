echo "== full GPU suite with VFM_PAIR_COLGROUPS=1"
VFM_PAIR_COLGROUPS=1 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== timing (default, then VFM_PAIR_COLGROUPS=1)"
for cfg in "--cin 640 --cout 512 --res 32" "--cin 512 --cout 256 --res 64" "--cin 256 --cout 128 --res 128"; do
  echo "-- $cfg"; timeout 60 python tools/conv_probe.py $cfg --up 2 --mode fused --iters 5; VFM_PAIR_COLGROUPS=1 timeout 60 python tools/conv_probe.py $cfg --up 2 --mode fused --iters 5
done
