#!/bin/bash
# Per-kernel SASS opcode histogram of libgonova_hift.so (what proves tcgen05 / TMA: UTCHMMA, UTMALDG, UTMASTG, LDTM, STTM).
# usage: tools/sass_histogram.sh > profiles/r02_sass_opcodes.txt      (CPU only: cuobjdump reads the built library)
LIB=${1:-gonova_tts_b200/lib/libgonova_hift.so}
cuobjdump -sass "$LIB" | awk '
  /Function : / { fn=$3; next }
  /^\s+\/\*[0-9a-f]+\*\/\s+/ {
    op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); split(op, a, "."); base=a[1];
    n[fn]++; c[fn, base]++; ops[base]=1
  }
  END {
    key="UTCHMMA UTCQMMA UTCBAR UTMALDG UTMASTG UTMAPF UTMACCTL LDTM STTM UTCATOMSWS SYNCS MUFU HMMA ELECT REDG RED LDGSTS";
    nk=split(key, K, " ");
    for (f in n) {
      line=sprintf("%-110s total %6d ", f, n[f]);
      for (i=1;i<=nk;i++) if (c[f, K[i]]) line=line sprintf(" %s=%d", K[i], c[f, K[i]]);
      print line
    }
  }' | sort
