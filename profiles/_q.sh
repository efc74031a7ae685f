python profiles/nms_phases.py > gpurun_out/q_nms_phases.log 2>&1; cat gpurun_out/q_nms_phases.log | grep -v Warn
SWEEP_REPS=2 python profiles/sweep_k1.py > gpurun_out/q_sweep_k1.log 2>&1; cat gpurun_out/q_sweep_k1.log
bash profiles/ncu_kernels.sh r2p
