"""Compare the SASS of the cost-volume kernels in two object files instruction by instruction (parameter-bank
offsets normalised): used to prove that adding an opt-in template variant leaves the shipped kernels untouched.
    python scripts/compare_sass.py old.o new.o"""
import re,subprocess,collections,sys
def load(obj):
    out=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    d=collections.OrderedDict(); name=None
    for line in out.split('\n'):
        m=re.search(r'Function : (\S+)',line)
        if m: name=m.group(1); d[name]=[]; continue
        m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',line)
        if m and name: d[name].append(re.sub(r'c\[0x0\]\[0x[0-9a-f]+\]','c[P]',m.group(2).strip()))
    return d
def key(n):
    m=re.search(r'cost_volume_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb(\d)ELb(\d)E(?:Lb(\d)E)?Li(\d)E',n)
    return m.groups() if m else n
old=load(sys.argv[1]); new=load(sys.argv[2])
oldk={}
for n,v in old.items():
    k=key(n); oldk[(k[:8]+k[9:]) if isinstance(k,tuple) else k]=v
for n,v in new.items():
    k=key(n)
    if isinstance(k,tuple):
        if k[8]=='1': print('STORE variant',k,len(v),'instr'); continue
        kk=k[:8]+k[9:]
    else: kk=k
    o=oldk.get(kk)
    print(kk,'same (%d instr)'%len(v) if o==v else 'DIFF old %s new %d'%(len(o) if o else None,len(v)))
