timeout 900 ncu --set full --clock-control none --import-source on -k regex:"backward_kernel|forward_kernel" -c 4 -o gpurun_out/r02_gent_w8 -f python tools/dbg_generic_profile.py > gpurun_out/ncu_gent_w8.log 2>&1
tail -5 gpurun_out/ncu_gent_w8.log
