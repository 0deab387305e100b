timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for w in c1 c2 c3 c5; do echo "== $w"; timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${w}_r01b.json 2>gpurun_out/b.err; tail -3 gpurun_out/b.err; python -c "
import json
j=json.load(open('gpurun_out/bench_${w}_r01b.json'))
print('value %.4g e2e %.4g ms/step %.3f trmm_frac %.3f whole %.3f err %.2e cpu %s launches %d'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'] or 0, j['roofline']['whole_step_frac_of_peak'], j['config']['sample_max_rel_err_vs_oracle'], j.get('cpu_baseline',{}).get('value'), j['gpu_launches']), j['roofline']['share_of_step'], j['clocks'])"; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
