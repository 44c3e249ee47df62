for n in 24576 32768 49152; do for t in warp lane; do
AB200_TUNE=$t timeout 300 python bench.py --steps 1 --warmup 1 --cells $n --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n $t', round(d['value']), round(d['ms_per_step']), d['config']['evolution_build'][:4], d['reach_cells'])"
done; done
