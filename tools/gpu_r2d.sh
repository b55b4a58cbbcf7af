mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log); tail -3 gpurun_out/gpu_tests.log
bash tools/gpu_ncu_others.sh
python tools/sanitize_run.py --prepare /tmp/states.npy --instances 600 --age 1500 > gpurun_out/sanitizer_prepare.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_run.py --run /tmp/states.npy --steps 2 > gpurun_out/sanitizer_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|ran [0-9]+ steps" gpurun_out/sanitizer_racecheck.log | tail -3
grep -o "in sim_kernel.cu:[0-9]*" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -20
