mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_block_kernel -c 1 -o gpurun_out/r03_jacobi_block -f python tools/small_probe.py c2 1 > gpurun_out/r03_ncu_jb.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel_kernel -s 3 -c 1 -o gpurun_out/r03_lu_panel_cl -f python tools/small_probe.py c2 1 > gpurun_out/r03_ncu_lp.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
