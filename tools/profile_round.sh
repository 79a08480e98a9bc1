#!/bin/bash
# Usage (on the GPU box, through gpurun): tools/profile_round.sh <out_dir under gpurun_out> 
# The measured evidence of a round: the bench lines of every configuration, the reference arm, ncu launch lists and one
# `ncu --set full` capture of a whole config-3 step (bring the .ncu-rep back and read it with tools/ncu_traffic.py / ncu_lines.py).
out=gpurun_out/$1
mkdir -p "$out"
python bench.py > "$out/bench_c3.json" 2> "$out/bench_c3.err"
for c in 1 2 4 5; do python bench.py --config $c > "$out/bench_c$c.json" 2> "$out/bench_c$c.err"; done
python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference_c3.json" 2> "$out/bench_reference_c3.err"
for c in 3 4 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_c$c.csv" python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-pipelining > "$out/ncu_launches_c$c.log" 2>&1
done
ncu --set full --clock-control none --import-source on -s 69 -c 24 -o "$out/step_c3" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pipelining > "$out/ncu_full_c3.log" 2>&1
cuobjdump -sass contrast_renderer_b200/libcontrast_b200.so | grep -B2 -A6 "UBLKCP" | head -60 > "$out/sass_ublkcp.txt"
