# round-2 (second session) ncu captures of the new default kernels + launch list of the bench command; one GPU
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
# TV-L1 persistent TMA kernel (multi-warp TMA issue, TMA-store epilogue), level 0, K = 8: launch 166 of 200
timeout 300 $NCU --kernel-name regex:k_tvl1_blocked_tma --launch-skip 165 -o gpurun_out/prof_tvl1_tma_r2u python tools/prof_one.py tvl1 8 1 0 > /dev/null 2>&1
# TV-L1 tiled warp kernel, level 0: launch 46 of 50
timeout 300 $NCU --kernel-name regex:k_tvl1_warp_tile --launch-skip 45 -o gpurun_out/prof_tvl1_warp_tile_r2u python tools/prof_one.py tvl1 8 1 0 > /dev/null 2>&1
ls -la gpurun_out/*r2u.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2u_tvl1.csv python bench.py --steps 1 --warmup 3 --pairs 2 --streams 1 --no-cpu --no-extras --workload tvl1 > gpurun_out/r2u_bench_under_ncu.log 2>&1
python tools/ncu_durations.py gpurun_out/launches_r2u_tvl1.csv
