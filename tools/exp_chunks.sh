# experiment: chunk size / slots of the GPU BAM decoder vs stream time (run on the GPU box)
python tools/make_bam.py /dev/shm/x.bam > /dev/null
for cfg in "48 8" "64 4" "64 6" "64 8" "96 4" "96 6" "128 4" "128 5" "192 3" "192 4"; do
  set -- $cfg
  for i in 1 2 3; do excord_lr_b200/host/excord-lr-b200 -b /dev/shm/x.bam -o /dev/shm/x.txt -t 16 --chunk-mb $1 --slots $2 --stats 2>&1 | grep -o "stream [0-9.]* s" | tr '\n' ' '; done; echo " mb=$1 slots=$2"
done
