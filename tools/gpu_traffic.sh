#!/usr/bin/env bash
# DRAM traffic and L2 hit rate of k_apg_persistent per setting of the L2 knobs (ncu metrics pass, second launch of a solve).
# Usage (under gpurun): bash tools/gpu_traffic.sh <tag> <workload> "<RN_L2_PREFETCH values>" [iters=10]
set -uo pipefail
TAG="${1:-traf}"; W="${2:-C2}"; IT="${4:-10}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum"
for pf in ${3:-0 0.4}; do
  RN_L2_PREFETCH=$pf timeout 300 ncu --metrics $M --clock-control none -k regex:k_apg_persistent --launch-skip 1 -c 1 --csv --log-file "$OUT/traffic_${W}_pf$pf.csv" python tools/ncu_solve.py $W $IT > "$OUT/traffic_${W}_pf$pf.log" 2>&1
  echo "ncu $W pf=$pf rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/traffic_${W}_pf$pf.csv" "$W" "$pf" "$IT" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
d={r[mi]:(r[vi],r[ui]) for r in rows[1:]}
def val(k):
    v,u=d[k]; v=float(v.replace(",",""))
    return v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
it=int(sys.argv[4])
print(f"TRAFFIC {sys.argv[2]} pf={sys.argv[3]}: read {val('dram__bytes_read.sum')/it/1e6:.1f} MB/iter, write {val('dram__bytes_write.sum')/it/1e6:.1f} MB/iter, L2 hit {d['lts__t_sector_hit_rate.pct'][0]} %, duration {d['gpu__time_duration.sum']}")
PY
done
