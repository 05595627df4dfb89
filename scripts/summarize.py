import json, sys, csv, subprocess
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'ERR', e); continue
    r=d['roofline']
    print(f, 'ms/step', round(d['ms_per_step'],2), 'pairs/s %.3g'%d['value'], 'frac', round(r['frac'],3), 'GB/s', round(r['achieved']), 'step ms avg', round(r['ms_per_launch_avg'],3), 'share', round(r['share_of_step'],3), 'expand ms', round(r['expand_ms'] or 0,2), 'expand GB/s', round(r['expand_gbs'] or 0), 'e2e', d['e2e'] and round(d['e2e']['seconds_per_step'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
