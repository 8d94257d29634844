timeout 600 python scripts/exp_halo_bound.py 400 > gpurun_out/r2_exp_halo_bound3.txt 2>&1; echo exp rc=$?; cat gpurun_out/r2_exp_halo_bound3.txt
