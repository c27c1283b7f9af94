for cfg in "0 0" "16 8" "32 16" "32 0" "48 24"; do set -- $cfg; 
  export MODE_TRAIN_COMM_SMS=$1; if [ "$2" != "0" ]; then export NCCL_MAX_CTAS=$2; else unset NCCL_MAX_CTAS; fi
  echo "COMM_SMS=$1 NCCL_MAX_CTAS=$2"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$1 scripts/train_bench.py 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
done
