set -x
mkdir -p gpurun_out/z
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/z/pytest.log; cat gpurun_out/z/pytest.log
python bench.py > gpurun_out/z/bench.json 2> gpurun_out/z/bench.err; tail -2 gpurun_out/z/bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/z/ref_arm.json 2> gpurun_out/z/ref_arm.err
for w in rach edge vitac wideband modulate; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/z/bench_$w.json 2> gpurun_out/z/bench_$w.err; done
for w in vitac wideband; do python bench.py --impl reference --workload $w --steps 3 --warmup 1 > gpurun_out/z/ref_arm_$w.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'corr_|peak_kernel|demod_kernel' -c 60 --csv --log-file gpurun_out/z/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/z/ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:demod_kernel -c 1 -o gpurun_out/z/demod python tools/prof_case.py --case nb --n 262144 > gpurun_out/z/d.log 2>&1
python tools/kernel_bench.py --json gpurun_out/z/kernel_rooflines.json > gpurun_out/z/kernel_rooflines.txt 2>&1; tail -5 gpurun_out/z/kernel_rooflines.txt
ls -la gpurun_out/z
