#!/bin/bash
# Static SASS statistics per kernel of a cubin: instructions, MUFU, control flow (BRA/BSSY/BSYNC/CALL/RET).
# usage: scripts/sass_count.sh file.cubin
cuobjdump -sass "$1" | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {cnt[name]++; if ($0 ~ /MUFU/) mu[name]++; if ($0 ~ /BRA|BSSY|BSYNC|CALL|RET|WARPSYNC/) br[name]++} END {for (n in cnt) printf "%6d inst %4d mufu %4d ctrl  %s\n", cnt[n], mu[n], br[n], n}' | c++filt | sed 's/(float const.*//; s/(mvae::PmParams)//' | sort -k7
