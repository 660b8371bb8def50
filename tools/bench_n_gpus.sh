# usage (on an N-GPU box): bash tools/bench_n_gpus.sh N [tag] [extra bench flags]
N=$1
TAG="${2:-run}"
shift; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${N}gpu.json 2> gpurun_out/${TAG}_${N}gpu.err
tail -3 gpurun_out/${TAG}_${N}gpu.err
timeout 20 python tools/bench_brief.py gpurun_out/${TAG}_${N}gpu.json
