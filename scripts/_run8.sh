N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 scripts/multi_sweep.py 16 64 256 1024 2>&1 | grep "ms per sweep"
NCCL_MAX_CTAS=8 $TR --master-port 29512 scripts/multi_sweep.py 64 2>&1 | grep "ms per sweep"
NCCL_MAX_CTAS=4 $TR --master-port 29513 scripts/multi_sweep.py 64 2>&1 | grep "ms per sweep"
NCCL_MIN_CTAS=32 $TR --master-port 29514 scripts/multi_sweep.py 64 2>&1 | grep "ms per sweep"
NCCL_ALGO=Ring $TR --master-port 29515 scripts/multi_sweep.py 64 2>&1 | grep "ms per sweep"
NCCL_ALGO=NVLS $TR --master-port 29516 scripts/multi_sweep.py 64 2>&1 | grep "ms per sweep"
