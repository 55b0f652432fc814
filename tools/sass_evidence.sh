#!/bin/bash
# Blackwell-native evidence from the built library (no GPU needed): for the main kernels, the SASS lines that show
# tcgen05 (UTC*MMA, LDTM/STTM), TMA / bulk copies (UBLKCP, UTMA*), mbarriers (SYNCS) and cp.async (LDGSTS), plus
# register / shared-memory usage.  Output: profiles/sass/*.txt
set -e
cd "$(dirname "$0")/.."
SO=transiflow_b200/lib/libtfb200.so
OUT=profiles/sass
mkdir -p $OUT
cuobjdump -res-usage $SO 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/^ *//' | grep -E "assemble_march_kernelI[0-9]Cfg_(ldc3d|rb3d)Lb1ELb1E|spmv_march_kernelI9Cfg_(ldc3d|rb3d)|tfb_fdm_plane_kernel|tfb_thomas_kernel|k_idr_sweep|k_shadow_dots|k_all_axpy|tfb_tc_pre4|tfb_tc_post4|k_gj_blocked" > $OUT/r2_resource_usage.txt
for pat in "tfb_fdm_plane_kernelILi32ELi3E" "tfb_assemble_march_kernelI9Cfg_ldc3dLb1ELb1ELi1ELi16ELi4ELi0ELi1E" "tfb_assemble_march_kernelI8Cfg_rb3dLb1ELb1ELi2ELi16ELi1ELi0ELi1E" "k_gj_blockedILi512E" "tfb_spmv_march_kernelI9Cfg_ldc3dLi2ELi16ELb0E"; do
    fn=$(cuobjdump -res-usage $SO 2>/dev/null | grep -o "Function [^:]*" | awk '{print $2}' | grep "$pat" | head -1)
    [ -z "$fn" ] && continue
    short=$(echo $pat | sed 's/I[0-9]Cfg_/_/; s/[^A-Za-z0-9_]/_/g')
    {
        echo "# cuobjdump -sass -fun $fn $SO  (filtered to the async-proxy / tensor-core / barrier instructions)"
        echo "# mnemonic histogram:"
        cuobjdump -sass -fun "$fn" $SO 2>/dev/null | grep -oE "\b(UTC[A-Z0-9]*MMA[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOM[.A-Z0-9_]*|LDTM[.A-Z0-9_]*|STTM[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|UTMA[A-Z]*[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|LDGSTS[.A-Z0-9_]*|BAR\.[A-Z._]*|DFMA|DMUL|DADD|FFMA|LDG[.A-Z0-9_]*|STG[.A-Z0-9_]*|LDS[.A-Z0-9_]*|STS[.A-Z0-9_]*|FENCE[.A-Z0-9_]*|MEMBAR[.A-Z0-9_]*)" | sort | uniq -c | sort -rn
        echo "# lines:"
        cuobjdump -sass -fun "$fn" $SO 2>/dev/null | grep -E "UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|LDGSTS|FENCE|ARRIVE" | sed 's/^ *//' | head -80
    } > $OUT/r2_sass_$short.txt
done
ls -la $OUT
