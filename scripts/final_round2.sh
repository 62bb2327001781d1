#!/bin/bash
# Round-end evidence on one B200: full -m gpu suite, the three bench lines + the reference arm, ncu (full set + launch list),
# compute-sanitizer over smoke().  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/final_pytest.log
timeout 400 python bench.py 2>/dev/null | tail -1 > gpurun_out/final_huf.json
timeout 400 python bench.py --codec fse 2>/dev/null | tail -1 > gpurun_out/final_fse.json
timeout 400 python bench.py --codec u16 2>/dev/null | tail -1 > gpurun_out/final_u16.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:huf_(plan|emit|decode)_kernel' -s 8 -c 4 -o gpurun_out/r02_huf_final -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck.log 2>&1
cat gpurun_out/final_pytest.log; tail -3 gpurun_out/san_memcheck.log; tail -3 gpurun_out/san_racecheck.log
for f in huf fse u16 ref; do python -c "
import json; d=json.load(open('gpurun_out/final_$f.json')); print('$f', d.get('value'), (d.get('e2e') or {}).get('value'), d.get('bit_exact'), (d.get('roofline') or {}).get('all_kernels'))"; done
