# A/B of bench.py argument sets on the same box: bash profiles/scripts/ab_args.sh "ARGS_A" "ARGS_B" ["ARGS_C" ...]
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "|", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], round(d["ms_per_step"],1))'
F="--steps 3 --warmup 3 --no-cpu-baseline --no-train --no-accuracy --no-library-baseline"
rm -f gpurun_out/ab.log
for r in 1 2; do
  for a in "$@"; do
    python bench.py $F $a 2>>gpurun_out/ab.err | python -c "$P" "$a" >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
