# A/B of one environment switch on the same box: bash profiles/scripts/ab_env.sh NAME VALUE_A VALUE_B [rounds]
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], round(d["ms_per_step"],1))'
F="--steps 3 --warmup 3 --no-cpu-baseline --no-train --no-accuracy --no-library-baseline"
rm -f gpurun_out/ab.log
for i in $(seq 1 ${4:-2}); do
  env $1=$2 python bench.py $F 2>>gpurun_out/ab.err | python -c "$P" "$1=$2" >> gpurun_out/ab.log
  env $1=$3 python bench.py $F 2>>gpurun_out/ab.err | python -c "$P" "$1=$3" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
