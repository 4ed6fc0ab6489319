python tools/make_bam.py /dev/shm/x.bam > /dev/null
for mb in 32 64 128 256 600; do
  for i in 1 2; do excord_lr_b200/host/excord-lr-b200 -b /dev/shm/x.bam -o /dev/shm/x.txt -t 16 --chunk-mb $mb --stats 2>&1 | grep "GPU BAM decoder" | sed "s/^/mb=$mb /" | cut -c1-330; done
done
