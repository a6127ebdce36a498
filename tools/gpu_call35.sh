#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_api_cuda.py tests/test_transfer_cuda.py tests/test_engine_cuda.py -x -q -k "api or transfer or em_ or estimate" > gpurun_out/c35_pytest.log 2>&1; echo "pytest: $?" | tee gpurun_out/c35.log
tail -n 3 gpurun_out/c35_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c35_bench_c3_$i.json 2> gpurun_out/c35_bench_c3_$i.err; echo "bench $i: $?" | tee -a gpurun_out/c35.log
done
python - <<'P'
import json
for f in ('c35_bench_c3_1','c35_bench_c3_2'):
    for l in open('gpurun_out/%s.json'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, '%.4g'%d['value'], d['ms_per_step'], 'e2e %.4g in %.3f s'%(d['e2e']['value'], d['e2e']['seconds']), d['e2e']['phases_s'])
P
tail -n 3 gpurun_out/c35_bench_c3_1.err
