# Round-2 evidence of the CURRENT build in one gpurun call (< 64 MiB back): smoke, the pairing capture bench.py's
# roofline.traffic is read from, the verify capture, the launch list of bench.py, the bench line itself.
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
P="python tools/prof_run.py"
NCU="ncu --set full --clock-control none"
$P pairing 56832 4 1 > gpurun_out/plain_pairing.log 2>&1 && $NCU --import-source on -k regex:vm_kernel -s 1 -c 1 -o gpurun_out/r2_pairing_fullwave $P pairing 56832 4 1 > gpurun_out/ncu1.log 2>&1
$P verify 56832 4 1 > gpurun_out/plain_verify.log 2>&1 && $NCU -k regex:vm_kernel -s 3 -c 1 -o gpurun_out/r2_verify_full $P verify 56832 4 1 > gpurun_out/ncu2.log 2>&1
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu8.log 2>&1
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>&1
cat gpurun_out/plain_*.log; tail -c 300 gpurun_out/r2_bench_1gpu.err
