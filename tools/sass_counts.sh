#!/bin/bash
# usage: tools/sass_counts.sh [lib.so]  ->  stdout: SASS mnemonic counts of the shipped library (profiles/rNN/sass_counts.txt)
lib=${1:-quantizedmha_b200/lib/libqmha.so}
sass=$(mktemp); cuobjdump -sass "$lib" > "$sass"
echo "cuobjdump -sass $lib | grep -c <mnemonic>   (library built by 'make lib' from this commit; whole-word"
echo "matches for the legacy tensor mnemonics, so UTCHMMA / UTCIMMA do not count as HMMA / IMMA)"
for m in UTCIMMA UTCHMMA UTMALDG UTMASTG LDTM STTM UTCBAR MUFU.EX2; do printf "%-12s %s\n" $m "$(grep -c "$m" "$sass")"; done
for m in HMMA IMMA HGMMA IGMMA WGMMA; do printf "%-12s %s   (legacy mma.sync / wgmma: must be 0)\n" $m "$(grep -cw "$m[.A-Z0-9]*" "$sass")"; done
echo; echo "kernels in the library (instantiations per kernel):"
grep "Function :" "$sass" | sed 's/.*Function : //' | c++filt | sed 's/void qmha::(anonymous namespace)::\([a-z_0-9]*\)<.*/\1/; s/void qmha::\([a-z_0-9]*\)<.*/\1/' | sort | uniq -c | sort -rn
rm -f "$sass"
