#!/bin/bash
# C5 on 2 GPUs (strong): value under the environment given on the command line.  usage: scripts/c5_probe2.sh TAG [VAR=value ...]
TAG=$1; shift
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus ${NGPU:-2} --workload c5 --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$TAG', '$*', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms', d.get('parity_check'), d['gpu_launches'])"
