#!/bin/bash
# profiles/sass_evidence.sh -- which tensor-core / TMA / TMEM instructions the policy-forward kernels of the shipped library
# contain (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mma.sync -> HMMA)
LIB=${1:-hhmarl_2d_b200/csrc/libhhmarl_b200.so}
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn = $3 }
  { for (i = 1; i <= NF; ++i) if ($i ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|HMMA|UTCBAR|UTCATOMSWS|SYNCS)(\.|$)/) { split($i, a, ";"); c[fn "  " a[1]]++ } }
  END { for (k in c) print c[k], k }' | sort -k2,2 -k3,3 | grep -i "policy_forward"
