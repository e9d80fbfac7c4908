set -x
mkdir -p gpurun_out/z3
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/z3/pytest.log; cat gpurun_out/z3/pytest.log
python bench.py > gpurun_out/z3/bench.json 2> gpurun_out/z3/bench.err; tail -2 gpurun_out/z3/bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/z3/ref_arm.json 2> gpurun_out/z3/ref_arm.err
for w in rach edge vitac wideband modulate; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/z3/bench_$w.json 2> gpurun_out/z3/bench_$w.err; done
for w in vitac wideband; do python bench.py --impl reference --workload $w --steps 3 --warmup 1 > gpurun_out/z3/ref_arm_$w.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'detect_lane|corr_|peak_kernel|demod_kernel' -c 60 --csv --log-file gpurun_out/z3/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/z3/ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:demod_kernel -c 1 -f -o gpurun_out/z3/demod python tools/prof_case.py --case nb --n 262144 > gpurun_out/z3/d.log 2>&1
python tools/kernel_bench.py --json gpurun_out/z3/kernel_rooflines.json > gpurun_out/z3/kernel_rooflines.txt 2>&1; tail -5 gpurun_out/z3/kernel_rooflines.txt
for c in nb; do ncu --set full --import-source on --clock-control none -k regex:detect_lane -c 1 -f -o gpurun_out/z3/detlane_$c python tools/prof_case.py --case $c --n 262144 > /dev/null 2>&1; done
ncu --set full --import-source on --clock-control none -k regex:demod_kernel -c 1 -f -o gpurun_out/z3/demod_edge python tools/prof_case.py --case edge --n 262144 > /dev/null 2>&1
ls -la gpurun_out/z3
