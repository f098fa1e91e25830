#!/bin/bash
# static SASS instruction count per kernel of an object / shared library:  scripts/sass_count.sh file
cuobjdump -sass "$1" | grep -E "Function :|^\s+/\*[0-9a-f]{4,}\*/" | awk '/Function :/ {name=$3; next} {cnt[name]++} END {for (k in cnt) print cnt[k], k}' | sort -rn
