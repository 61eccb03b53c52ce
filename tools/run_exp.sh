timeout 300 python tools/profile_step.py 256 2 2>&1 | tail -1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"bottleneck|conv_wgrad" --launch-skip 11 -c 11 -o gpurun_out/prof_wg python tools/profile_step.py 256 2 > gpurun_out/ncu_wg.log 2>&1; tail -1 gpurun_out/ncu_wg.log
python tools/ncu_summary.py gpurun_out/prof_wg.ncu-rep "Round 2 (final): weight-gradient GEMMs and the fused bottleneck kernels of one eager training step at batch 256 (ncu --set full)" > gpurun_out/r02_wgrad_bottleneck_full.md
python tools/ncu_traffic.py gpurun_out/prof_wg.ncu-rep "x" | tail -8
rm -f gpurun_out/prof_wg.ncu-rep
