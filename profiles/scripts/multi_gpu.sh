#!/bin/bash
# bench.py on N GPUs of one box the way the driver launches it (strong scaling of the named batch); N=$1, extra args follow
N=$1; shift
TAG=${TAG:-r02}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/bench_${TAG}_n$N$SUFFIX.json 2> gpurun_out/bench_${TAG}_n$N$SUFFIX.err
tail -1 gpurun_out/bench_${TAG}_n$N$SUFFIX.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('frac_of_h2d_ceiling'), d['clocks'].get('reasons'))"
