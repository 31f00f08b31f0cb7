#!/usr/bin/env bash
# instruction count / bytes of every function of the persistent kernel (the sweeps are bound by instruction fetch)
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
T=$(mktemp -d); cd "$T"
cuobjdump -xelf all "$ROOT/rapidnet_b200/librapidnet_b200.so" >/dev/null
nvdisasm -c rn_persist.sm_100a.cubin > cur.sass
python - <<'PY'
import re
cur=None; cnt={}
for ln in open('cur.sass'):
    m=re.match(r'^(\$_Z\S+|\.text\.\S+):',ln)
    if m: cur=m.group(1); cnt[cur]=0; continue
    if cur and re.match(r'^\s+/\*[0-9a-f]+\*/',ln): cnt[cur]+=1
tot=0
for k,v in sorted(cnt.items(), key=lambda x:-x[1]):
    if 'gemv_roleILi' in k and 'Li3ELi1E' not in k: continue
    name=re.sub(r'^.*\$_ZN2rn\d+','',k)[:40]
    print(f"{v:6d} {v*16/1024:7.1f} KB  {name}"); tot+=v
print("total (one gemv instance)", tot, f"{tot*16/1024:.1f} KB")
PY
rm -rf "$T"
