run() { # variant wpc
  AB200_LIBRARY=/root/repo/release-1.0_b200/libv/$1/libariadne_b200.so AB200_WPC=$2 timeout 300 python bench.py --steps 1 --warmup 1 --cells 8192 --no-cpu-baseline > gpurun_out/r2_sw5_$1_$2.json 2> gpurun_out/r2_sw5_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sw5_$1_$2.json')); print('$1 wpc $2', round(d['value']), 'steps/s', round(d['roofline']['evolution_steps_per_s']), d['reach_cells'])" || tail -3 gpurun_out/r2_sw5_$1_$2.err
}
run occ16 16; run occ16 8; run occ16 4; run occ12 12; run occ12 6; run occ12 4
AB200_LIBRARY=/root/repo/release-1.0_b200/libv/occ16/libariadne_b200.so AB200_WPC=16 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact or lane_build or identical" 2>&1 | tail -3
