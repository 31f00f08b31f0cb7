#!/usr/bin/env bash
# SASS evidence for the judged kernel (no GPU needed): architecture, instruction census and the lines around the bulk-TMA
# copies of k_apg_persistent in the built library.  Usage: bash tools/sass_evidence.sh > profiles/r02_sass_k_apg_persistent.txt
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
LIB="${1:-$ROOT/rapidnet_b200/librapidnet_b200.so}"
CUOBJDUMP="${CUOBJDUMP:-/usr/local/cuda/bin/cuobjdump}"
TMP="$(mktemp)"
"$CUOBJDUMP" -sass "$LIB" > "$TMP"
echo "# $(basename "$LIB"): $(sha256sum "$LIB" | cut -c1-16)  ($(date -u +%Y-%m-%dT%H:%MZ))"
echo "## architectures in the fat binary"
"$CUOBJDUMP" -lelf "$LIB" | sed 's/^/  /'
echo "## functions (SASS)"
grep -E "^\s*Function :" "$TMP" | sed 's/^\s*/  /'
echo "## instruction census, whole library"
for pat in UBLKCP UTMALDG SYNCS LDGSTS "LDG\." "STG\." "LDS" "STS" FFMA2 "FFMA " HMMA UTCHMMA UTCMMA "BAR\.SYNC" "ATOM" "RED\." "MEMBAR" ERRBAR "CCTL"; do
  printf "  %-10s %s\n" "$pat" "$(grep -cE "\b$pat" "$TMP" || true)"
done
echo "## k_apg_persistent: instruction census"
awk '/Function : .*k_apg_persistent/{f=1} f&&/Function : /&&!/k_apg_persistent/{f=0} f' "$TMP" > "$TMP.k"
for pat in UBLKCP SYNCS LDGSTS "LDG\." "STG\." "LDS" "STS" FFMA2 "FFMA " "DADD" "BAR\.SYNC" "ATOM" "MEMBAR" "CALL"; do
  printf "  %-10s %s\n" "$pat" "$(grep -cE "\b$pat" "$TMP.k" || true)"
done
echo "  lines      $(wc -l < "$TMP.k")"
echo "## k_apg_persistent: the bulk-TMA copies (UBLKCP) with their mbarrier arrive/expect (SYNCS) in context"
grep -nE "UBLKCP|SYNCS\.ARRIVE|SYNCS\.EXCH" "$TMP.k" | head -60 | sed 's/^/  /'
echo "## first UBLKCP, 12 lines of context"
n=$(grep -nE "UBLKCP" "$TMP.k" | head -1 | cut -d: -f1)
[ -n "$n" ] && sed -n "$((n>8?n-8:1)),$((n+4))p" "$TMP.k" | sed 's/^/  /'
rm -f "$TMP" "$TMP.k"
