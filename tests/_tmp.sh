for f in 1 0; do RTGR_VERBOSE=1 RTGR_CHUNK_RAYS=$f python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02y_err$f.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('CHUNK_RAYS=$f 4K: kernel path %.2f ms; e2e %.2f ms (kernel %s); pageable %.2f ms' % (d['kernel_ms_per_step'], e['ms_per_step'], e['kernel_ms_per_rank'], e['pageable_host_buffer']['ms_per_step']))"; grep "rtgr:" gpurun_out/r02y_err$f.log | head -3; done
