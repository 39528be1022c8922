#!/bin/bash
# usage: smoke_loop.sh <mask> <count>
fails=0
for i in $(seq 1 $2); do
  out=$(MOONSHINE_B200_DEBUG_SYNC=$1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "moonshine-b200\|smoke ok" | tail -2 | tr '\n' ' ')
  case "$out" in *"smoke ok"*"rvkaka"*) ;; *) fails=$((fails+1)); echo "mask $1 iter $i: $out";; esac
done
echo "mask $1: $fails failures / $2"
