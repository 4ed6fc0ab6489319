# The GPU BAM decoder across the GPUs of one box: one synthetic HiFi BAM (configs[1] molecules x $1 with 15 kb SEQ/QUAL per record),
# chunks round robin over --gpus N, steady-state stream figure (--eager-alloc), best of 3 per N.  Output identical across N (cmp).
#   gpurun --gpus 2 -- 'bash tools/bam_scaling.sh 0.12 "1 2" > gpurun_out/bam_scaling.txt 2>&1'
scale=${1:-0.12}; ns=${2:-"1 2"}
python tools/make_bam.py /dev/shm/big.bam --scale $scale
for n in $ns; do
  for rep in 1 2 3; do
    excord_lr_b200/host/excord-lr-b200 -b /dev/shm/big.bam -o /dev/shm/out_$n.txt -p 0.8 -t 32 --gpus $n --eager-alloc --stats 2>&1 | grep -o "stream [0-9.]* s\|[0-9]* chunks, [0-9.]* MB of BGZF\|on [0-9]* of [0-9]* GPU" | tr '\n' ' '
    echo " (--gpus $n)"
  done
  cmp /dev/shm/out_1.txt /dev/shm/out_$n.txt && echo "output of --gpus $n identical to --gpus 1"
done
