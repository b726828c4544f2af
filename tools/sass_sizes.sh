#!/bin/bash
# instruction count per kernel / device function in the product library (code size matters: the single-lane EPA kernel is
# instruction fetch bound). usage: tools/sass_sizes.sh [lib.so]
LIB=${1:-joltphysics_b200/libjolt_b200.so}
cuobjdump -sass "$LIB" | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | sort -rn | c++filt | cut -c1-180 | head -${2:-25}
