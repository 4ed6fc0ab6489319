# experiment: GPU BAM decoder on N GPUs -- batches in flight per GPU and chunk size (2.3 GB BAM, steady-state stream figure)
n=${1:-2}
python tools/make_bam.py /dev/shm/big.bam --scale 0.12 > /dev/null
for mb in 64 128; do for sl in 2 3 4 5 7; do
  for rep in 1 2; do
    excord_lr_b200/host/excord-lr-b200 -b /dev/shm/big.bam -o /dev/shm/o.txt -p 0.8 -t 32 --gpus $n --chunk-mb $mb --slots $sl --eager-alloc --stats 2>&1 | grep -o "stream [0-9.]* s\|record walk [0-9.]* s\|walk + gather [0-9.]* ms\|inflate [0-9.]* ms" | tr '\n' ' '
  done
  echo " mb=$mb slots=$sl gpus=$n"
done; done
