import re,sys,subprocess,collections
obj=sys.argv[1]
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
for f in re.split(r'\n\s*Function : ',txt)[1:]:
    name=f.split('\n',1)[0]
    if 'tcgemm_kernelIf' not in name: continue
    ins=[]
    for l in f.split('\n'):
        m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2)))
    best=None
    for a,t in ins:
        if 'BRA' in t:
            tg=re.search(r'0x([0-9a-f]+)',t)
            if tg and int(tg.group(1),16)<a:
                lo=int(tg.group(1),16)
                body=[x for b,x in ins if lo<=b<=a]
                has=lambda k: any(k in x for x in body)
                if has('LDGSTS') and has('STS.128') and has('SYNCS.ARRIVE') and len(body)<1500:
                    if best is None or len(body)<best[0]: best=(len(body),lo,a)
    tag=re.search(r'tcgemm_kernelIf(Li\dELi\dELi\dELi\dELi\dE)',name).group(1)
    print(tag, 'producer main loop', best)
