#!/bin/bash
# One-GPU acceptance visit: what the driver runs at round end (GPU tests, smoke, both bench arms with the driver's flags),
# plus the default bench command.
OUT=gpurun_out
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/gputests_final.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/gputests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench_final_ref.json 2> $OUT/bench_final_ref.err; echo "ref rc=$?"; tail -1 $OUT/bench_final_ref.json | cut -c1-400; tail -3 $OUT/bench_final_ref.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench_final_driver.json 2> $OUT/bench_final_driver.err; echo "bench(driver flags) rc=$?"; tail -4 $OUT/bench_final_driver.err
( time timeout 900 python bench.py ) > $OUT/bench_final_default.json 2> $OUT/bench_final_default.err; echo "bench(default) rc=$?"; tail -4 $OUT/bench_final_default.err
python - <<PY
import json
for f in ('driver','default'):
    d=json.loads(open('$OUT/bench_final_%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value %.3fM ms %.4f e2e %.3fM frac %.4f f32 %.3fM r1w %s cpu %s'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['roofline']['frac'],d['config']['e2e_float32_observations']['value']/1e6,d['config'].get('same_kernel_on_round1_env_starts'),d.get('cpu_baseline')))
    for s in d['secondary']: print('   ',s['workload'][:44],'value %.3fM e2e %.3fM ms %.3f frac %.4f'%(s['value']/1e6,s['e2e']/1e6,s['ms_per_step'],s['roofline_frac']))
PY
