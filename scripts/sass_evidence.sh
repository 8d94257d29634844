#!/bin/sh
# Count the Blackwell-specific SASS instructions per kernel of the built library (runs without a GPU):
#   UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
#   SYNCS = mbarrier ops, REDG = global reductions, UCGABAR = cluster barrier.
cuobjdump -sass "${1:-evfly_b200/libevfly_b200.so}" 2>/dev/null | awk '
/Function :/ {fn=$3}
{ for (i = 1; i <= NF; ++i) if ($i ~ /^(UTCHMMA|UTMALDG|LDTM|UTCBAR|REDG|SYNCS|UCGABAR)/) { op=$i; sub(/\..*/, "", op); c[fn" "op]++ } }
END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{n=$NF; op=$(NF-1); $NF=""; $(NF-1)=""; sub(/\(.*/, "", $0); gsub(/^void /, "", $0); gsub(/evfly::/, "", $0); a[$0]=a[$0] " " op "=" n} END {for (k in a) print k ":" a[k]}' | sort
