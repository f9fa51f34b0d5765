"""Summary of a tools/tc_timeline.py log: mean cycles of conversion, arrive->issue lag, issue, drain, period per stage."""
import re, sys, statistics as S
for fn in sys.argv[1:]:
    rows = {}
    for l in open(fn).read().splitlines():
        m = re.match(r'it(\d) st(\d+) slot(\d) (.*)', l)
        if not m: continue
        d = {k: (int(v) if v != '-' else None) for k, v in re.findall(r'(\w+)=\s*(-|\d+)', m.group(4))}
        rows[(int(m.group(1)), int(m.group(2)), int(m.group(3)))] = d
    conv, lag, iss, drain, per = [], [], [], [], []
    for (it, st, sl), d in rows.items():
        if None in (d.get('done'), d.get('stored'), d.get('arrived'), d.get('ready'), d.get('committed')): continue
        conv.append(d['stored'] - d['done']); lag.append(d['ready'] - d['arrived']); iss.append(d['committed'] - d['ready'])
        nx = rows.get((it, st + 1, sl))
        if nx and nx.get('done') is not None:
            drain.append(nx['done'] - d['committed']); per.append(nx['done'] - d['done'])
    f = lambda a: '%5.0f' % S.mean(a) if a else '    -'
    print('%-40s conv %s  lag %s  issue %s  drain %s  period %s' % (fn.split('/')[-1], f(conv), f(lag), f(iss), f(drain), f(per)))
