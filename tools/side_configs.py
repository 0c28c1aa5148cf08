import json,sys
d=json.loads(open(sys.argv[1]).readline())
print(sys.argv[1], d['ms_per_step'], d['config']['rank0_ms'])
for k,v in d['extra']['configs'].items(): print(' ',k, 'k3',round(v['k3_ms'],3),'k4',round(v['k4_ms'],3), 'step', round(v['ms_per_step'],3))
