set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02k_tests_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1
tail -n 5 gpurun_out/r02k_tests_all.log gpurun_out/r02k_smoke.log
