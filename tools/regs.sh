#!/bin/bash
# registers / spills of the three variant step kernels in a build log
for f in "$@"; do
  echo "== $f"
  awk '/Compiling entry function.*k_stepILi(4|9|16)ELi(4|9|16)ELb[01]ELb0/{name=$0} /Used/{ if(name!=""){ match(name,/k_stepILi[0-9]+ELi[0-9]+ELb[01]/); print substr(name,RSTART,RLENGTH), $0; name=""} } /spill/{ if ($5+0>0 || $9+0>0) print "   SPILL", $0 }' "$f"
done
