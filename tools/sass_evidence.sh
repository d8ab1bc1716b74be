#!/bin/bash
# profiles/r2_sass_evidence.txt: counts of the mnemonics that prove the 1-D bulk TMA / mbarrier / cp.async paths, per kernel
echo "# SASS evidence (cuobjdump -sass elynx_b200/libelynx_b200.so, sm_100a cubins; built by elynx_b200/build.py; tools/sass_evidence.sh)"
echo "# counts of the mnemonics that prove the 1-D bulk TMA / mbarrier / cp.async paths, per kernel"
cuobjdump -sass elynx_b200/libelynx_b200.so | awk '
/Function : /{fn=$3}
/\/\*[0-9a-f]+\*\//{ for(i=1;i<=NF;i++) if($i ~ /^\/\*[0-9a-f]+\*\/$/){op=$(i+1); if(op ~ /^@/) op=$(i+2); sub(/;/,"",op); break}
  if (op ~ /^(UBLKCP|SYNCS|LDGSTS|IMAD\.WIDE\.U32|PRMT|VIMNMX|STG\.E)/) c[fn" "op]++ }
END{ for(k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { print "== "$1; last=$1 } printf "%7d %s\n", $3, $2 }' | awk '
/^== /{ keep = ($0 ~ /k2_quad_tiled|k2_duo_tiled|k2_tiled|k2_pair_tiled|k_duo_unpermute_async/) } keep{print}'
