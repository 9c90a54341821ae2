#!/bin/bash
# SASS evidence for the tensor-core / TMA / barrier instructions of every kernel family (cuobjdump on the built objects).
# usage: tools/sass_summary.sh > profiles/r02f_sass_summary.txt
B=flow_field_based_motion_planner_b200/csrc/build
for o in qnet step flow_field_il flow_field_wide flow_field flow_field_large host_io replay; do
  echo "== $o.o"
  cuobjdump -sass $B/$o.o 2>/dev/null | grep -oE "Function : [A-Za-z0-9_]+|UTC[A-Z0-9]+(\.[A-Z0-9_]+)*|LDTM(\.[A-Z0-9x]+)*|UTMALDG(\.[A-Z0-9]+)*|UBLKCP(\.[A-Z0-9]+)*|SYNCS(\.[A-Z0-9_]+)*|ACQBULK|UCGABAR_[A-Z]+|STAS(\.[0-9]+)?|MAPA|PRMT|SHFL(\.[A-Z]+)*|REDUX(\.[A-Z]+)*|IMAD(\.[A-Z0-9]+)*|LOP3(\.LUT)?|SHF(\.[A-Z0-9]+)*|BAR(\.[A-Z0-9]+)*" \
    | awk '/Function/{f=$3; next}{c[f"\t"$0]++}END{for(k in c)print k"\t"c[k]}' | sort | c++filt | awk -F'\t' '{printf "%-28s %6d  %s\n", $2, $3, substr($1,1,110)}' | grep -vE "^(IMAD|LOP3|SHF|PRMT|BAR|SHFL)" 
  echo "-- logic / shuffle / barrier instruction counts per kernel (static)"
  cuobjdump -sass $B/$o.o 2>/dev/null | grep -oE "Function : [A-Za-z0-9_]+|PRMT|SHFL|IMAD|LOP3|SHF\.|BAR\.SYNC|BAR\.RED" \
    | awk '/Function/{f=$3; next}{c[f"\t"$0]++}END{for(k in c)print k"\t"c[k]}' | sort | c++filt | awk -F'\t' '{printf "%-10s %6d  %s\n", $2, $3, substr($1,1,110)}'
done
