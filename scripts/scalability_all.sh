# round-2 scalability sweep (BASELINE config 4): arcs 5M / 20M / 50M, k = 500, on N GPUs; usage: scalability_all.sh N out.csv
N=$1; OUT=$2
if [ "$N" = "1" ]; then
  timeout 900 python scripts/scalability.py --arcs 5000000 20000000 50000000 --k 500 > $OUT 2> $OUT.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N scripts/scalability.py --arcs 5000000 20000000 50000000 --k 500 > $OUT 2> $OUT.err
fi
cat $OUT
