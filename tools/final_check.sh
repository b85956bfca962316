set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu_r1y.txt
python bench.py > gpurun_out/bench_r1y.json 2> gpurun_out/bench_r1y.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1y.json 2> gpurun_out/bench_ref_r1y.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1y.txt 2>&1; echo "smoke rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1y.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r1y.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/pytest_gpu_r1y.txt; cut -c1-600 gpurun_out/bench_r1y.json; cut -c1-700 gpurun_out/bench_ref_r1y.json; tail -3 gpurun_out/smoke_r1y.txt
