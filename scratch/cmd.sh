timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "wgrad_rdb" 2>&1 | tail -2
for S in 4 6 8; do
B200_WGRAD_RDB_KSPLIT=$S B200_BENCH_DETAIL=gpurun_out/detail_s$S.txt python bench.py --steps 10 --warmup 5 --no-cudnn-ref --no-cpu-baseline > gpurun_out/b_s$S.json 2>gpurun_out/b.err
python -c "
import json; d=json.load(open('gpurun_out/b_s$S.json')); print('S=$S', d['ms_per_step'], d['e2e']['ms_per_step'])"
grep "^wgrad_rdb" gpurun_out/detail_s$S.txt
done
