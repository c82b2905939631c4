#!/usr/bin/env python
"""Opcode mix of the kernels in an ncu report: share of executed warp instructions and (in brackets) of stall samples per SASS opcode.

    ncu -i gpurun_out/x.ncu-rep --page source --csv --print-source sass > /tmp/x.csv
    python tools/ncu_opcode_mix.py /tmp/x.csv
"""
import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
sections=[];cur=None;i=0
while i<len(rows):
    r=rows[i]
    if r and r[0]=='Kernel Name':
        cur={'func':r[1],'hdr':rows[i+1],'rows':[]}; sections.append(cur); i+=2; continue
    if cur is not None and r: cur['rows'].append(r)
    i+=1
for s in sections:
    hdr=s['hdr']; ie=hdr.index('Instructions Executed'); src=hdr.index('Source'); sm=hdr.index('# Samples')
    tot=0; cls=collections.Counter(); smp=collections.Counter(); ts=0
    for r in s['rows']:
        try: v=int(r[ie]); sa=int(r[sm])
        except: continue
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)',r[src])
        op=m.group(2).split('.')[0] if m else '?'
        cls[op]+=v; tot+=v; smp[op]+=sa; ts+=sa
    print('=====',s['func'][40:185],'total',tot,'static',len(s['rows']))
    print('  '.join('%s:%.1f%%(%.0f%%)'%(k,100*v/tot,100*smp[k]/max(ts,1)) for k,v in cls.most_common(30)))
