# round-2 final measurement pass on one B200 (outputs in gpurun_out/, copied to profiles/ afterwards)
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -n 2 gpurun_out/r2_gpu_tests.log
python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_ref.err
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_b.err; tail -n 2 gpurun_out/r2_b.err
rm -f gpurun_out/r2_bench_configs.jsonl
for c in 0 1 2 3 4; do python bench.py --config $c --no-e2e-bam 2>/dev/null | tail -n 1 >> gpurun_out/r2_bench_configs.jsonl; done
python tools/cli_timing.py > gpurun_out/r2_cli_timing.txt 2>&1; grep "best of 3" gpurun_out/r2_cli_timing.txt | cut -c1-220
python tools/timeline.py 1 0 q > gpurun_out/r2_timeline.txt 2>&1
bash tools/r2_profile.sh
