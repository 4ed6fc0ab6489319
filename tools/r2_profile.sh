# round-2 profiling pass on the GPU box (B200_PROFILING.md recipe): launch list of the bench command, one --set full capture of
# every kernel of a step, and of the BAM decoder's kernels.  Outputs land in gpurun_out/; profiles/summarize.py digests them.
python tools/make_bam.py /dev/shm/x.bam > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e-bam --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k[0-9]' --launch-skip 16 -c 8 -f -o gpurun_out/r2_all_kernels python tools/profile_step.py --graph 0 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kb_ -c 8 -f -o gpurun_out/r2_bam_kernels excord_lr_b200/host/excord-lr-b200 -b /dev/shm/x.bam -o /dev/shm/x.txt --chunk-mb 600 > gpurun_out/ncu_bam.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log gpurun_out/ncu_bam.log
