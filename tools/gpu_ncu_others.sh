# ncu summaries of the kernels besides the fused k_step: k_reset, k_obs, k_ram, k_pack, k_order, k_flags (configs[1] size)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'k_reset|k_obs|k_ram|k_pack|k_order|k_flags' -c 40 -o gpurun_out/others_full -f python tools/exercise_kernels.py 4096 > gpurun_out/ncu_others.log 2>&1
tail -2 gpurun_out/ncu_others.log | cut -c1-200
