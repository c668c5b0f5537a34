mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel_kernel -s 2 -c 12 -o gpurun_out/r03_lu_panel_cl3 -f python tools/small_probe.py c2 1 > gpurun_out/r03_ncu_lp.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*cl3.ncu-rep
