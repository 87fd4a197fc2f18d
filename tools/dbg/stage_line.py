import sys, json
d = json.loads(sys.stdin.readline())
st = {s['stage']: round(s['ms_per_step'], 2) for s in d['stages']}
print(sys.argv[1], round(d['value']), round(d['e2e']['value']), {k: st[k] for k in ('candidates', 'resolve', 'describe', 'fast', 'line_merge', 'lsd_grow')})
