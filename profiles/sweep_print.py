import json, glob, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/sw_*.json')):
    try:
        d = json.load(open(f)); r = d['roofline']
        print(f.split('/')[-1], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'p50 %.3f' % d['p50_align_ms'], 'kern us %.0f' % r['avg_launch_us'],
              'pairs/pt %.1f' % r['pairs_per_point'], 'frac %.3f' % (r['frac'] or 0), 'e2e %.0f' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e)
