#!/bin/bash
# usage: tools/sass_loop_hist.sh <lib.so> <kernel-name-substring>
# Prints the opcode histogram of the innermost back-edge loop that contains MUFU.SIN
# (the per-epoch loop of the likelihood kernel).
so=$1; pat=$2
cuobjdump -sass "$so" | awk -v pat="$pat" '/Function : /{f=index($0,pat)>0} f' > /tmp/_k.sass
python3 - << 'PY'
import re
lines=[l for l in open('/tmp/_k.sass') if re.match(r'\s+/\*[0-9a-f]{4,6}\*/',l)]
ins=[]
for l in lines:
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
# find backward branches
best=None
for a,t in ins:
    m=re.search(r'BRA\s+(?:P\d,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a:
            body=[x for x in ins if tgt<=x[0]<=a]
            if any('MUFU.SIN' in x[1] for x in body):
                if best is None or len(body)<len(best): best=body
from collections import Counter
c=Counter()
for a,t in best:
    t=re.sub(r'^@!?U?P\d\s+','',t)
    c[t.split()[0].split('.')[0]]+=1
fp64=sum(v for k,v in c.items() if k in('DFMA','DMUL','DADD','DSETP'))
print('loop 0x%x-0x%x: %d instrs, %d FP64'%(best[0][0],best[-1][0],len(best),fp64))
print(sorted(c.items(),key=lambda x:-x[1]))
PY
