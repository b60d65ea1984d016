#!/usr/bin/env bash
# ncu evidence for the update -> recovery fusion (opt-in, measured slower): one --set full capture of k_rk_prim
set -u
O=gpurun_out/r02g; mkdir -p $O
XF_FUSE_PRIM=1 ncu --set full --clock-control none -k 'regex:k_rk_prim|k_prim_shell' -s 4 -c 2 -o $O/r02_full_rkprim -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --profile-steps 0 > $O/ncu.log 2>&1
python tools/ncu_summary.py full $O/r02_full_rkprim.ncu-rep > $O/r02_ncu_rk_prim.md 2>/dev/null
rm -f $O/*.ncu-rep
grep -E "^## |duration|regs/thread|achieved occ|issue slots|FP64 pipe|dram read|dram write|DRAM throughput|top stall|occ. limit" $O/r02_ncu_rk_prim.md | cut -c1-220
