#!/bin/bash
# usage: sass_hist.sh <object> <function-substring>  -> opcode histogram of the matching function(s)
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{f=$3} /\/\*[0-9a-f]+\*\/ +[A-Z@]/{ if (f ~ pat) { op=$2; if (op ~ /^@/) op=$3; sub(/;.*/,"",op); c[op]++; n++ } } END{for (k in c) print c[k], k; print n, "TOTAL"}' | sort -rn
