#!/usr/bin/env python3
"""print the per-kernel roofline table of bench.py JSON lines (dev tool)"""
import json
import sys
for f in sys.argv[1:]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001
        print(f, 'ERR', e)
        continue
    r = l['roofline']
    print(f, 'value %.1f it/s  ms/step %.4f fused=%s variant=%s launches=%s' % (l['value'], l['ms_per_step'], l['config'].get('fused_updates'), l['config']['spmv_variant'], l['gpu_launches']))
    for k in r['kernels']:
        print('   slot %d %-50s %.4f ms  %.0f GB/s frac %.3f share %.3f (n=%d)' % (k['slot'], k['kernel'][:50], k['avg_launch_ms'], k['GBps'], k['frac'], k['share_of_step'], k['launches_timed']))
    print('   iteration', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r['iteration'].items()})
    if 'csr_kernel' in r:
        c = r['csr_kernel']
        print('   csr %.1f it/s, iteration frac %.3f (of 8TB/s %.3f)' % (c['value'], c['iteration_frac'], c['iteration_frac_of_8TBps']))
    print('   clocks', l['clocks'], 'converge', l['converge'].get('iterations'), l['converge'].get('bit_identical_to_cpu_oracle'))
    if 'e2e' in l:
        print('   e2e', json.dumps(l['e2e'])[:700])
    for k in ('reference_gpu_ilu0', 'ilu0', 'ilu0_multicolor', 'mat10000_ilu0', 'cpu_baseline', 'random_dd_50M', 'cusparse_spmv'):
        if k in l:
            print('  ', k, json.dumps(l[k])[:500])
    if 'poisson512' in l['config']:
        print('   poisson512', json.dumps(l['config']['poisson512'])[:400])
