import csv, sys
f=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 60
rows=list(csv.reader(open(f)))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[0]!="Address"]
tot=sum(int(r[idx['# Samples']]) for r in data)
totinst=sum(int(r[idx['Instructions Executed']]) for r in data)
print('total samples',tot,'total warp-instr',totinst)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={s:sum(int(r[idx[s]]) for r in data) for s in stalls}
print({k:v for k,v in sorted(agg.items(), key=lambda x:-x[1]) if v>0})
# opcode mix
import collections
mix=collections.Counter()
for r in data:
    op=r[idx['Source']].split()[0]
    if op.startswith('@'): op=r[idx['Source']].split()[1]
    mix[op.split('.')[0]]+=int(r[idx['Instructions Executed']])
print([(k,round(100*v/totinst,1)) for k,v in mix.most_common(25)])
print('--- top by samples')
for i,r in sorted(enumerate(data), key=lambda x:-int(x[1][idx['# Samples']]))[:top]:
    st={s:int(r[idx[s]]) for s in stalls if int(r[idx[s]])>0}
    st=sorted(st.items(), key=lambda x:-x[1])[:3]
    print(f"{i:5d} {int(r[idx['# Samples']]):6d} {int(r[idx['Instructions Executed']]):9d} {r[idx['Source']].strip()[:70]:70s} {st}")
