set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu_r1z.txt
python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1z.json 2> gpurun_out/bench_ref_r1z.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1z.txt 2>&1; echo "smoke rc=$?"
cat gpurun_out/pytest_gpu_r1z.txt; cut -c1-600 gpurun_out/bench_r1z.json; cut -c1-700 gpurun_out/bench_ref_r1z.json; tail -3 gpurun_out/smoke_r1z.txt
