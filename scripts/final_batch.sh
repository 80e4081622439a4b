set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.txt; cat gpurun_out/final_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; cut -c1-300 gpurun_out/final_bench_reference.json
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; cut -c1-200 gpurun_out/final_bench_n1.json
python scripts/profile_target.py 4 50000000 2 > gpurun_out/plain_tiled50M.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tiled -s 2 -c 2 -o gpurun_out/prof_tiled50M_ws python scripts/profile_target.py 4 50000000 2 > gpurun_out/ncu_tiled50M.log 2>&1; tail -2 gpurun_out/ncu_tiled50M.log
