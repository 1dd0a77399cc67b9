# development container, after tools/evidence_round2.sh came back: summaries into profiles/ (build ids from THIS tree)
python tools/ncu_summary.py gpurun_out/r2_pairing_fullwave.ncu-rep profiles/r2_pairing "ncu --set full --clock-control none: vm_kernel<true,1,384> (one thread per item, CTA-wide item blocks), program pairing@4, round-2 final build, n = 56,832 = exactly one full wave" > /dev/null
python tools/ncu_summary.py gpurun_out/r2_verify_full.ncu-rep profiles/r2_verify_full "ncu --set full --clock-control none, round-2 final build: verify_full@4 (hash-to-G2 + dual Miller loop + cubed final exponentiation), vm_kernel<true,1,384>, n = 56,832 = one full wave" > /dev/null
python - <<'PY'
import json
for src, dst in (("gpurun_out/r2_bench_1gpu.json", "profiles/r2_bench_1gpu.json"), ("gpurun_out/r2_bench_reference_arm.json", "profiles/r2_bench_reference_arm.json")):
    line = [l for l in open(src) if l.startswith("{")][-1]
    json.dump(json.loads(line), open(dst, "w"), indent=1)
PY
cp gpurun_out/r2_bench_launches.csv profiles/r2_bench_launches.csv
grep -E "duration|dram__bytes|fmaheavy|registers" profiles/r2_pairing_ncu_summary.txt profiles/r2_verify_full_ncu_summary.txt
python -c "
import json; d=json.load(open('profiles/r2_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['extra']['verify_signatures_per_s'])"
