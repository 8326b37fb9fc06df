#!/bin/bash
# Per-kernel counts of the Blackwell-specific SASS instructions in libtokenpacker_b200.so (evidence that the hot kernels are
# tcgen05 / TMEM / TMA code):  tools/sass_counts.sh > profiles/r02_sass_counts.txt      (runs without a GPU)
lib=${1:-tokenpacker_b200/libtokenpacker_b200.so}
echo "# cuobjdump -sass $lib | per-function instruction counts"
echo "# UTCHMMA[.2CTA] = tcgen05.mma (cta_group::1 / ::2); LDTM = tcgen05.ld; UTCATOMSWS = tcgen05.alloc; UTCBAR = tcgen05.commit;"
echo "# UTMALDG = TMA tensor load; UTMASTG = TMA tensor store; UTMAPF = TMA L2 prefetch; SYNCS = mbarrier; FFMA2/FMUL2/FADD2 = packed fp32"
cuobjdump -sass "$lib" | awk '
  /Function :/ { fn=$3; order[++n]=fn; next }
  { for (i in pat) if ($0 ~ pat[i]) c[fn,i]++ }
  BEGIN { split("UTCHMMA.2CTA UTCHMMA LDTM UTCATOMSWS UTCBAR UTMALDG UTMASTG UTMAPF SYNCS FFMA2 FMUL2 FADD2 MUFU REDG LDG.E.STRONG", names, " ");
          for (i in names) pat[i]=names[i]; pat[2]="UTCHMMA[^.]" }
  END { for (k=1;k<=n;k++) { fn=order[k]; line=""; tot=0; for (i=1;i<=15;i++) { v=c[fn,i]+0; if (v>0) { line=line sprintf(" %s=%d", names[i], v); tot+=v } }
        if (tot>0) printf "%s\n   %s\n", fn, line } }' | c++filt
