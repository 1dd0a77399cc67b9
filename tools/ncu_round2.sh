# ncu captures of the round-2 build; gpurun brings back at most 64 MiB per call, hence two parts: bash tools/ncu_round2.sh A|B
# (each ncu run is preceded by the plain run of the same command)
set -x
P="python tools/prof_run.py"
NCU="ncu --set full --clock-control none"
if [ "$1" = "A" ]; then
$P pairing 56832 4 1 > gpurun_out/plain_pairing.log 2>&1 && $NCU --import-source on -k regex:vm_kernel -s 1 -c 1 -o gpurun_out/r2_pairing_fullwave $P pairing 56832 4 1 > gpurun_out/ncu1.log 2>&1
$P verify 56832 4 1 > gpurun_out/plain_verify.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 1 -o gpurun_out/r2_verify_full $P verify 56832 4 1 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
python bench.py --steps 4 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu8.log 2>&1
else
$P g2_sum 1000000 0 1 > gpurun_out/plain_g2sum.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 2 -o gpurun_out/r2_g2_sum $P g2_sum 1000000 0 1 > gpurun_out/ncu3.log 2>&1
$P g1_sum 1000000 0 1 > gpurun_out/plain_g1sum.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 2 -o gpurun_out/r2_g1_sum $P g1_sum 1000000 0 1 > gpurun_out/ncu4.log 2>&1
$P g2_msm 1000000 0 1 > gpurun_out/plain_g2msm.log 2>&1 && $NCU -k regex:vm_kernel -s 5 -c 4 -o gpurun_out/r2_g2_msm $P g2_msm 1000000 0 1 > gpurun_out/ncu5.log 2>&1
B200BLS_KERNEL=2 $P pairing 9472 0 1 > gpurun_out/plain_k2.log 2>&1 && B200BLS_KERNEL=2 $NCU -k regex:vm2_kernel -s 1 -c 1 -o gpurun_out/r2_pairing_paired_lowlat $P pairing 9472 0 1 > gpurun_out/ncu7.log 2>&1
fi
cat gpurun_out/plain_*.log
ls -la gpurun_out/
