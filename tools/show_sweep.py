import json, sys
for fn in sys.argv[1:]:
    print(fn)
    for l in open(fn):
        if not l.startswith('{'): print(l.strip()[:200]); continue
        d=json.loads(l)
        if 'error' in d: print(d); continue
        print("n=%-9d nq=%-5d k=%-3d used=%-14s call %.4f scan %.4f roof %.4f  frac_call %.3f frac_scan %.3f hits/q %-7.1f other %s %s" % (d['n'],d['nq'],d['k'],d['used'],d['call_ms'],d['scan_ms'],d['roof_ms'],d['frac_call'],d['frac_scan'],d['hits_per_q'],d['other_us'], d.get('cycles_pct','')))
