# ncu captures of the aggregation kernels of the final round-2 build (each preceded by the plain run of the same command)
set -x
P="python tools/prof_run.py"
NCU="ncu --set full --clock-control none"
$P g2_sum 1000000 0 1 > gpurun_out/plain_g2sum.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 2 -o gpurun_out/r2_g2_sum $P g2_sum 1000000 0 1 > gpurun_out/ncu3.log 2>&1
$P g1_sum 1000000 0 1 > gpurun_out/plain_g1sum.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 2 -o gpurun_out/r2_g1_sum $P g1_sum 1000000 0 1 > gpurun_out/ncu4.log 2>&1
# MSM: VM launches per call = tomont, bucket fold, bscale, sum1, sum2 (5); skip the first call's five
$P g2_msm 1000000 0 1 > gpurun_out/plain_g2msm.log 2>&1 && $NCU -k regex:vm_kernel -s 5 -c 3 -o gpurun_out/r2_g2_msm $P g2_msm 1000000 0 1 > gpurun_out/ncu5.log 2>&1
$P pairing 75776 5 1 > gpurun_out/plain_shape5.log 2>&1 && $NCU -k regex:vm_kernel -s 1 -c 1 -o gpurun_out/r2_pairing_shape5 $P pairing 75776 5 1 > gpurun_out/ncu6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_msm_launches.csv $P g2_msm 1000000 0 1 > gpurun_out/ncu7.log 2>&1
cat gpurun_out/plain_*.log
