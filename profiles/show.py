"""print the bench lines of a gpurun log written by the call scripts (cfg=... line followed by a JSON line)"""
import json, sys
cfg = None
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith('cfg='):
        cfg = l[4:]; continue
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(f"{cfg:28s} ms={d['ms_per_step']:9.2f} passes={d['config']['passes']:3d} rounds={d['config']['rounds']:4d} frac={r['frac']:.3f} gates/s={d['value']:.0f}")
    elif l:
        print(cfg, l[:300])
