# round-1 final state (r1_m): full GPU suite incl. the SURVEY 8(f) rows, smoke, bench (both arms)
set -x
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/r1m_pytest_gpu.log 2>&1
tail -25 gpurun_out/r1m_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1m_smoke.log 2>&1
tail -1 gpurun_out/r1m_smoke.log
timeout 280 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/r1m_step_profile.txt > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err
tail -c 1500 gpurun_out/r1m_bench.json
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1m_bench_reference.json 2> gpurun_out/r1m_bench_reference.err
