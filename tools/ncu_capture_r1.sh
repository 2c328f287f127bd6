set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
# TV-L1 persistent TMA kernel, level 0, K = 8 (one calc = 200 iteration launches, level 0 = the last 40)
$NCU --kernel-name regex:k_tvl1_blocked_tma --launch-skip 165 -o gpurun_out/prof_tvl1_tma_r1d python tools/prof_one.py tvl1 8 1 > /dev/null 2>&1
# Brox register-resident SOR, level 0 of 720p (280 launches, level 0 = the last 20)
$NCU --kernel-name regex:k_brox_sor_reg --launch-skip 265 -o gpurun_out/prof_brox_sor_r1a python tools/prof_brox.py 720 1280 > /dev/null 2>&1
# DensePyrLK fast kernel, level 0 (4 launches, the last)
$NCU --kernel-name regex:k_lk_dense_fast --launch-skip 3 -o gpurun_out/prof_lk_fast_r1a python tools/prof_one.py denselk 0 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
