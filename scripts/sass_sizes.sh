#!/bin/bash
# sizes (bytes of SASS) of the sub-functions of a kernel in libpmdi_cuda.so: scripts/sass_sizes.sh k_sweep_pool
K=${1:-k_sweep_pool}
D=$(mktemp -d); cd $D
cuobjdump -xelf all ${PMDI_LIB:-/root/repo/particlemdi.jl_b200/libpmdi_cuda.so} >/dev/null 2>&1
nvdisasm -g -c *.cubin > /tmp/all_g.sass 2>/dev/null
python3 - "$K" <<'PY'
import re, sys
K=sys.argv[1]
lines=open('/tmp/all_g.sass').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.%s:'%K)][0]
end=[i for i,l in enumerate(lines) if i>start and l.startswith('//--------------------- .text.')]
end=end[0] if end else len(lines)
addr=0; labs=[('main',0)]
for l in lines[start:end]:
    m2=re.search(r'/\*([0-9a-f]{4,6})\*/',l)
    if m2: addr=int(m2.group(1),16)
    if l.startswith('$'): labs.append((l.strip().split('$')[2][:48],addr))
labs.append(('end',addr))
for i in range(len(labs)-1): print(f'{labs[i][0]:50s} {labs[i+1][1]-labs[i][1]:7d}')
print('total', addr)
PY
rm -rf $D
