import re,sys,collections,subprocess
obj,pat=sys.argv[1],sys.argv[2]
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
funcs=re.split(r'\n\s*Function : ',txt)
for f in funcs[1:]:
    name=f.split('\n',1)[0]
    if pat not in name: continue
    ins=[]
    for l in f.split('\n'):
        m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2)))
    print(name[:150]); print('  total SASS',len(ins))
    for a,t in ins:
        if 'BRA' in t:
            tg=re.search(r'0x([0-9a-f]+)',t)
            if tg and int(tg.group(1),16)<a and (a-int(tg.group(1),16))//16>20:
                lo=int(tg.group(1),16)
                body=[re.sub(r'^@!?U?P\d+\s+','',x) for b,x in ins if lo<=b<=a]
                c=collections.Counter(x.split()[0].split('.')[0] for x in body)
                print('  loop %#x..%#x: %d instr  %s'%(lo,a,len(body),dict(c.most_common(12))))
