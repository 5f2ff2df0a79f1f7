#!/bin/bash
# tools/sass_hot.sh <FunctorName> [lib] -- static count of the hot path of k_map_flat<Functor, U, 0>: instructions between
# the first global load and the last global store of the all-vector tile, per element, plus registers / spills.
F=$1; LIB=${2:-numericals_b200/lib/libnuk.so}
cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o "_ZN3nuk10k_map_flatINS_[0-9]*${F}ELi[0-9]*ELi0EE[A-Za-z0-9_]*" | sort -u | head -1)" $LIB 2>/dev/null > /tmp/_hot.sass
python3 - "$F" <<'P'
import re,sys
L=[l for l in open('/tmp/_hot.sass') if re.match(r'\s+/\*[0-9a-f]{4,}\*/',l)]
ops=[re.sub(r'^\s*/\*[0-9a-f]+\*/\s*','',l).split(';')[0].strip() for l in L]
name=[l for l in open('/tmp/_hot.sass') if 'Function' in l]
ld=[i for i,o in enumerate(ops) if 'LDG' in o]; st=[i for i,o in enumerate(ops) if 'STG' in o]
if not ld or not st: print(sys.argv[1],'no kernel found'); sys.exit()
# the all-vector tile: first run of loads up to the store that precedes the next load-run
first=ld[0]
nl=[i for i in ld if i<st[0]]
nst=[i for i in st if not any(st[0]<j<i for j in ld)]
last=nst[-1]
m=re.search(r'ELi(\d+)ELi0',name[0]); U=int(m.group(1))
packs=len(nl); 
body=ops[first:last+1]
import collections
c=collections.Counter((o.split()[1] if o.startswith('@') else o.split()[0]).split('.')[0] for o in body)
w128=any('.128' in ops[i] for i in nl)
elems=packs*(4 if w128 else 2)
arity=1
print("%-8s unroll %d: %d loads, %d stores, %d instr in tile = %.1f / elem (assuming %d elems/thread); total kernel %d; top: %s" % (sys.argv[1],U,packs,len(nst),len(body),len(body)/elems,elems,len(ops)," ".join("%s:%d"%kv for kv in c.most_common(9))))
P
