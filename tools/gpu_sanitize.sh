# compute-sanitizer on k_step at steady state (600 instances at game age 1500): memcheck, racecheck, synccheck
mkdir -p gpurun_out
python tools/sanitize_run.py --prepare /tmp/states.npy --instances 600 --age 1500 > gpurun_out/sanitizer_prepare.log 2>&1; tail -1 gpurun_out/sanitizer_prepare.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py --run /tmp/states.npy --steps 2 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ran [0-9]+ steps" gpurun_out/sanitizer_$tool.log | tail -3
done
