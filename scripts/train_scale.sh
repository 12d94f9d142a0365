# Training step of the hot path at cfg 5's per-GPU shape on 1..N GPUs (weak scaling, SyncBatchNorm + DDP over NCCL),
# for each arithmetic of the forward / data-gradient convolutions.  usage: bash scripts/train_scale.sh <out dir> <max gpus>
OUT=${1:-gpurun_out/train}; MAXN=${2:-1}
mkdir -p $OUT
for MODE in ${MODES:-fp32 tf32}; do
  for N in 1 2 4 8; do
    [ $N -gt $MAXN ] && continue
    MVS_TRAIN_CONV=$MODE timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
      scripts/train_ddp_bench.py --steps 10 --warmup 3 2> $OUT/err_${MODE}_$N.log | grep '^{' | tee $OUT/train_${MODE}_$N.json | cut -c1-400
  done
done
