mkdir -p gpurun_out
set -x
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], d["ms_per_step"])'
F="--steps 3 --warmup 3 --no-cpu-baseline --no-train --no-accuracy --no-library-baseline"
for i in 1 2; do
APTAI_SERPENTINE=1 python bench.py $F 2>gpurun_out/ab.err | python -c "$P" serp1 >> gpurun_out/ab.log
APTAI_SERPENTINE=0 python bench.py $F 2>>gpurun_out/ab.err | python -c "$P" serp0 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
