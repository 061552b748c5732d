python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python tools/igemm_timeline.py 0 > gpurun_out/igemm_timeline.txt 2>&1
python tools/time_igemm.py > gpurun_out/igemm_layers.txt 2>&1
B200_IGEMM_DBG=1 python tools/time_igemm.py > gpurun_out/igemm_layers_noepi.txt 2>&1
B200_BENCH_DETAIL=gpurun_out/detail_final.txt python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['vs_cudnn'])"
