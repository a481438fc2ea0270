#!/bin/bash
# SASS excerpt of a spectra kernel: the TMA issue (UBLKCP), the mbarrier wait (SYNCS) and the first DFMA-dense stretch of the
# evaluation loop.  Usage: tools/sass_excerpt.sh [shift|kernel] > profiles/r2_sass_<...>.txt
#   shift  (default): cf_shift_kernel<M_LIN14, 7, 3, 3>, the default of df_mode 1 / 2 on 3+1D tiles
#   kernel          : cf_kernel<M_LIN14, 7, 3, false, 3, 4>, the round-1 default (still the kernel of df_mode 3 / 4, mode 2, 2+1D)
cd "$(dirname "$0")/.."
if [ "${1:-shift}" = kernel ]; then PAT='9cf_kernelILi1ELi7ELi3ELb0ELi3ELi4EEEvNS_9HotParamsE'; else PAT='cf_shift_kernelILi1ELi7ELi3ELi3EEEvNS_9HotParamsE'; fi
F=$(cuobjdump -sass is3d_b200/libis3d_b200.so 2>/dev/null | grep "Function : " | grep "$PAT" | head -1 | awk '{print $3}')
cuobjdump -sass -fun $F is3d_b200/libis3d_b200.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+//; s/\s+\/\* 0x[0-9a-f]+ \*\/\s*$//' > /tmp/sass_default.txt
echo "# cuobjdump -sass -fun $F is3d_b200/libis3d_b200.so   (sm_100a, $(wc -l < /tmp/sass_default.txt) instructions)"
echo "# opcode histogram of the whole kernel:"
awk '{for(i=2;i<=NF;i++) if ($i !~ /^@/) {print $i; break}}' /tmp/sass_default.txt | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -24 | awk '{printf "#   %6d %s\n", $1, $2}'
echo "# ---- TMA producer: cp.async.bulk (UBLKCP) + mbarrier arrive/expect-tx (SYNCS)"
grep -n "UBLKCP" /tmp/sass_default.txt | awk -F: 'NR==1 || $1 > last + 12 {print $1} {last = $1}' | head -2 | while read l; do sed -n "$((l-6)),$((l+8))p" /tmp/sass_default.txt; echo "..."; done
echo "# ---- consumer: mbarrier try-wait on the stage"
l=$(grep -n "SYNCS.PHASECHK" /tmp/sass_default.txt | head -1 | cut -d: -f1); [ -n "$l" ] && sed -n "$((l-2)),$((l+4))p" /tmp/sass_default.txt
echo "# ---- first DFMA-dense stretch of the evaluation loop (x, p.dsigma, exponential polynomial of a 3-member group)"
l=$(grep -n "DFMA" /tmp/sass_default.txt | awk -F: 'NR>40{print $1; exit}')
sed -n "${l},$((l+90))p" /tmp/sass_default.txt
