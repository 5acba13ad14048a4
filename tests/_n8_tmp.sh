mkdir -p gpurun_out/r02w
run() { # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02w/err_$label.log | grep "^{" | tail -1 > gpurun_out/r02w/bench_n8_$label.json
}
run default X=1
run rgb_to_device RTGR_DEBUG_RGB_TO_DEVICE=1
run chunk RTGR_CHUNK_RAYS=1
