#!/usr/bin/env bash
# usage: tools/sass_fn.sh <function-substring> [grep-pattern]  -- SASS of one (device) function of the persistent kernel's cubin
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
T=$(mktemp -d); cd "$T"
cuobjdump -xelf all "$ROOT/rapidnet_b200/librapidnet_b200.so" >/dev/null
nvdisasm -c rn_persist.sm_100a.cubin > all.sass
n=$(grep -n "^\$.*$1.*:\|^\.text\..*$1.*:" all.sass | head -1 | cut -d: -f1)
e=$(awk -v n="$n" 'NR>n && (/^\$_Z.*:$/ || /^\.text\./ || /\.type/) {print NR; exit}' all.sass)
sed -n "${n},${e}p" all.sass | grep -E "${2:-.}" | cut -c1-110
rm -rf "$T"
