#!/bin/bash
# Instruction census of the built library: proves which tensor / copy paths the kernels use (B200_PROFILING.md's SASS mnemonics).
#   bash tools/sass_summary.sh > profiles/sass_summary_rNN.txt
LIB=imagegenerator_b200/libsgb200.so
T=$(mktemp)
cuobjdump -sass $LIB | c++filt > $T
echo "cuobjdump -sass $LIB   ($(stat -c %s $LIB) bytes, $(grep -c 'Function :' $T) kernels, archs: $(cuobjdump -lelf $LIB | sed 's/.*\.\(sm_[0-9a-z]*\)\..*/\1/' | sort -u | tr '\n' ' '))"
echo
echo "whole library:"
for m in UTCHMMA "UTCHMMA.2CTA" UTCBAR LDTM UTMALDG "UTMALDG.*MULTICAST" UTMASTG HMMA "SYNCS" "REDG.E.ADD.F32" "REDG.E.ADD.F64" UTCATOMSWS; do
  printf "  %-22s %6d\n" "$m" "$(grep -c -E "\b$m" $T)"
done
echo
echo "per kernel (tcgen05 MMA / TMEM load / TMA load / legacy HMMA / instructions):"
printf "  %6s %6s %6s %6s %8s  %s\n" UTCHMMA LDTM UTMALDG HMMA instrs kernel
awk '/Function :/ {if (name != "" && u+l+t+h > 0) printf "  %6d %6d %6d %6d %8d  %s\n", u, l, t, h, n, name; name=$0; sub(/.*Function : /, "", name); name=substr(name,1,110); u=l=t=h=n=0}
     /UTCHMMA/ {u++} /LDTM/ {l++} /UTMALDG/ {t++} /HMMA/ && !/UTCHMMA/ {h++} /^ +\/\*[0-9a-f]+\*\// {n++}
     END {if (u+l+t+h > 0) printf "  %6d %6d %6d %6d %8d  %s\n", u, l, t, h, n, name}' $T
rm -f $T
