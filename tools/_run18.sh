for w in c1 c2 c3 c5; do timeout 900 python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/final_${w}.json 2>gpurun_out/b.err; tail -2 gpurun_out/b.err; python -c "
import json
j=json.load(open('gpurun_out/final_${w}.json'))
print('$w value %.4g e2e %.4g ms/step %.3f trmm_frac %.3f whole %.3f err %.2e cpu %.4g launches %d'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'] or 0, j['roofline']['whole_step_frac_of_peak'], j['config']['sample_max_rel_err_vs_oracle'], j.get('cpu_baseline',{}).get('value',0), j['gpu_launches']), j['clocks'])"; done
timeout 900 python bench.py --workload nll --steps 4 --warmup 3 --nll-n 4096 > gpurun_out/final_nll.json 2>gpurun_out/b.err; tail -2 gpurun_out/b.err; python -c "
import json
j=json.load(open('gpurun_out/final_nll.json'))
print('nll value %.1f e2e %.1f ms/step %.2f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac']), j.get('cpu_baseline'), j['clocks'])"
timeout 600 python bench.py --workload nll --steps 4 --warmup 3 --nll-n 2048 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print('nll2048 value %.1f e2e %.1f ms/step %.2f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac']), j.get('cpu_baseline'))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref_c2.json 2>/dev/null; cut -c1-260 gpurun_out/final_ref_c2.json
