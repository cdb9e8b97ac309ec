import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print("value %.3f e2e %.3f it/s, %.3f s/call" % (d['value'], d['e2e']['value'], d['e2e']['seconds_per_call']))
