mkdir -p gpurun_out/final
S=$(date +%s)
python -m pytest tests -x -q -m gpu > gpurun_out/final/gpu_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s)-S )) s"; S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s)-S )) s"; S=$(date +%s)
python bench.py --impl reference > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$? $(( $(date +%s)-S )) s"; S=$(date +%s)
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$? $(( $(date +%s)-S )) s"
tail -2 gpurun_out/final/gpu_tests.log; tail -c 300 gpurun_out/final/smoke.log; head -c 400 gpurun_out/final/bench.json
