timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for ch in 0 18944 37888 151552; do timeout 600 python bench.py --workload c2 --steps 8 --warmup 3 --no-cpu-baseline --chunk $ch 2>gpurun_out/b.err | python -c "
import sys,json
j=json.loads(sys.stdin.read())
print('chunk', j['config']['chunk'], 'value %.4g e2e %.4g ms/step %.3f trmm_frac %.3f whole %.3f err %.2e'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac'] or 0, j['roofline']['whole_step_frac_of_peak'], j['config']['sample_max_rel_err_vs_oracle']), j['roofline']['share_of_step'])"; tail -2 gpurun_out/b.err; done
