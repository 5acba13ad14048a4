mkdir -p gpurun_out/r02z
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_parity.py -m gpu -q -x -k "frame or multi_device or chunkwise" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02z/err_$label.log | grep "^{" | tail -1 > gpurun_out/r02z/bench_n8_$label.json
}
run default X=1
run raybyray RTGR_CHUNK_RAYS=0
timeout 300 python tests/multi_device_render.py 2>&1 | tail -6 | tee gpurun_out/r02z/multi_device_render.log
