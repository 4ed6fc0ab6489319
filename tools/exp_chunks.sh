# experiment: chunk size of the GPU BAM decoder vs stream time (run on the GPU box)
python tools/make_bam.py /dev/shm/x.bam > /dev/null
for mb in 32 64 96 128 256; do
  for i in 1 2; do excord_lr_b200/host/excord-lr-b200 -b /dev/shm/x.bam -o /dev/shm/x.txt -t 16 --chunk-mb $mb --stats 2>&1 | grep -o "[0-9]* chunks\|pinned memory [0-9.]* s\|record walk [0-9.]* s\|inflate [0-9.]* ms\|stream [0-9.]* s" | tr '\n' ';' | sed "s/^/mb=$mb /"; echo; done
done
