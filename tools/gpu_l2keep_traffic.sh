#!/usr/bin/env bash
# DRAM traffic, L2 hit rate and duration of k_apg_persistent per RN_L2_KEEP share (ncu metrics pass on the second launch of a solve).
# Usage (under gpurun): bash tools/gpu_l2keep_traffic.sh <tag> <workload> "<shares>" [iters=10]
set -uo pipefail
TAG="${1:-l2k}"; W="${2:-C2}"; IT="${4:-10}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum"
for k in ${3:-0 0.1 0.3 0.6}; do
  RN_L2_KEEP=$k timeout 40 ncu --metrics $M --clock-control none -k regex:k_apg_persistent --launch-skip 1 -c 1 --csv --log-file "$OUT/traffic_${W}_keep$k.csv" python tools/ncu_solve.py $W $IT > "$OUT/traffic_${W}_keep$k.log" 2>&1
  python - "$OUT/traffic_${W}_keep$k.csv" "$W" "$k" "$IT" <<'PY' | tee -a "$OUT/summary.txt"
import csv,sys
try:
    rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
    hdr=rows[0]; mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
    d={r[mi]:(r[vi],r[ui]) for r in rows[1:]}
    def val(k):
        v,u=d[k]; v=float(v.replace(",",""))
        return v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
    it=int(sys.argv[4])
    print(f"L2KEEP {sys.argv[2]} share={sys.argv[3]}: DRAM read {val('dram__bytes_read.sum')/it/1e6:.1f} MB/iter, write {val('dram__bytes_write.sum')/it/1e6:.1f} MB/iter, L2 sector hit rate {d['lts__t_sector_hit_rate.pct'][0]} %, duration {d['gpu__time_duration.sum'][0]} {d['gpu__time_duration.sum'][1]} for {it} iterations (under ncu)")
except Exception as ex:
    print("L2KEEP", sys.argv[3], "failed:", ex)
PY
done
